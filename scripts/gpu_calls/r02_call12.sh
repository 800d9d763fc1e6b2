#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
grep -E "worst|passed|failed|rc=|Error|error|FAILED" gpurun_out/r02_pytest_gpu.txt | tail -30
{
SE3TN_GRAPH=0 timeout 120 python scripts/trace_timeline.py bf16x3 1
SE3TN_GRAPH=0 timeout 120 python scripts/trace_timeline.py bf16x3 4
SE3TN_GRAPH=0 timeout 120 python scripts/trace_timeline.py bf16x3 64
} > gpurun_out/r02_trace4.txt 2>&1
cat gpurun_out/r02_trace4.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_g.json').read().strip().splitlines()[-1])
print('bench g:', d['value'], d['ms_per_step'], d.get('graph_launches_per_step'), d['single_track'])
P
