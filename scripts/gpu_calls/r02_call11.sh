#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
grep -E "worst|passed|failed|rc=|Error|error|FAILED" gpurun_out/r02_pytest_gpu.txt | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for m in 1 0; do SE3TN_GRAPH=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_f_graph$m.json 2> gpurun_out/r02_bench_f_graph$m.err; tail -3 gpurun_out/r02_bench_f_graph$m.err; done
python - <<'P'
import json
for m in (1,0):
    d=json.loads(open('gpurun_out/r02_bench_f_graph%d.json'%m).read().strip().splitlines()[-1])
    print('graph', m, d['value'], d['ms_per_step'], d.get('graph_launches_per_step'), d['e2e']['value'], d['single_track'])
P
