#!/bin/bash
# after batching the split-K partial loads: latency-mode tests, tracker/pyrender test, latency breakdown, trunk unit timeline at n = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "latency or pyrender or track_host or on_track or second_weight" 2>&1 | tail -15 > gpurun_out/r02_pytest_call27.txt
cat gpurun_out/r02_pytest_call27.txt
timeout 600 python scripts/latency_breakdown.py > gpurun_out/r02_latency_breakdown3.txt 2>&1
head -8 gpurun_out/r02_latency_breakdown3.txt | cut -c1-200
SE3TN_TRACE=1 timeout 300 python scripts/trunk_units.py 1 > gpurun_out/r02_trunk_units3.txt 2>&1
head -12 gpurun_out/r02_trunk_units3.txt | cut -c1-220
