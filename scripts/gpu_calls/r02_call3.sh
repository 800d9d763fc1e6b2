#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
tail -25 gpurun_out/r02_pytest_gpu.txt
{
timeout 120 python scripts/trace_timeline.py bf16x3 64
timeout 120 python scripts/trace_timeline.py bf16x3 1
timeout 120 python scripts/trace_timeline.py bf16 64
timeout 120 python scripts/trace_timeline.py tf32 64
} > gpurun_out/r02_trace2.txt 2>&1
cat gpurun_out/r02_trace2.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 2500 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
