#!/bin/bash
# se3tn_track_host: tests that go through the numpy path + latency breakdown
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "track_host or on_track or headless or ycb or tracker or drop_in or module_api" 2>&1 | tail -15 > gpurun_out/r02_pytest_track_host.txt
cat gpurun_out/r02_pytest_track_host.txt
timeout 600 python scripts/latency_breakdown.py > gpurun_out/r02_latency_breakdown2.txt 2>&1
head -30 gpurun_out/r02_latency_breakdown2.txt | cut -c1-200
