#!/bin/bash
# latency mode with per-piece block ownership: full GPU tests, latency breakdown, unit timeline at n = 1, driver-style bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_call29.txt
cat gpurun_out/r02_pytest_call29.txt
timeout 600 python scripts/latency_breakdown.py > gpurun_out/r02_latency_breakdown4.txt 2>&1
head -8 gpurun_out/r02_latency_breakdown4.txt | cut -c1-200
SE3TN_TRACE=1 timeout 300 python scripts/trunk_units.py bf16x3 1 > gpurun_out/r02_trunk_units4.txt 2>&1
head -9 gpurun_out/r02_trunk_units4.txt | cut -c1-230
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_call29.json 2> gpurun_out/r02_bench_call29.err
tail -c 600 gpurun_out/r02_bench_call29.json
