#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "track_host or on_track or headless or tracker" 2>&1 | tail -6
timeout 600 python scripts/latency_breakdown.py > gpurun_out/r02_latency_breakdown5.txt 2>&1
sed -n 1,9p gpurun_out/r02_latency_breakdown5.txt | cut -c1-200
