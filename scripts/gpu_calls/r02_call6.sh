#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/stem_ws_check.py > gpurun_out/r02_stem_ws.txt 2>&1; cat gpurun_out/r02_stem_ws.txt
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
grep -E "worst|passed|failed|rc=|Error|error|FAILED" gpurun_out/r02_pytest_gpu.txt | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_d.json').read().strip().splitlines()[-1])
print('bench d:', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms'], d['single_track'])
P
SE3TN_STEM_WS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_d_ws.json 2> gpurun_out/r02_bench_d_ws.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_d_ws.json').read().strip().splitlines()[-1])
print('bench d (stem ws):', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms'])
P
tail -3 gpurun_out/r02_bench_d_ws.err
