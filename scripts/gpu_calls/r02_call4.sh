#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
grep -E "worst|passed|failed|rc=|Error|error" gpurun_out/r02_pytest_gpu.txt | tail -30
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -c 3000 gpurun_out/r02_bench_b.json; tail -5 gpurun_out/r02_bench_b.err
