#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
grep -E "worst|passed|failed|rc=|Error|error|FAILED" gpurun_out/r02_pytest_gpu.txt | tail -30
{
timeout 120 python scripts/trace_timeline.py bf16x3 1
timeout 120 python scripts/trace_timeline.py bf16x3 8
timeout 120 python scripts/trace_timeline.py bf16x3 16
} > gpurun_out/r02_trace3.txt 2>&1
cat gpurun_out/r02_trace3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -c 1500 gpurun_out/r02_bench_c.json; tail -5 gpurun_out/r02_bench_c.err
SE3TN_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_c_nograph.json 2> gpurun_out/r02_bench_c_nograph.err
tail -c 1200 gpurun_out/r02_bench_c_nograph.json
timeout 60 ./scripts/umma_tmemA_probe > gpurun_out/r02_tmemA_probe.txt 2>&1; cat gpurun_out/r02_tmemA_probe.txt
