#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "on_track_end or tracker_with_cuda or headless or ycb_video" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_h.json').read().strip().splitlines()[-1])
print('bench h:', d['value'], d['ms_per_step'], d.get('graph_launches_per_step'), d['single_track'], d['e2e']['value'], d['weight_sets_21']['ratio_vs_1_set'], {k:v['value'] for k,v in d['alt_precisions'].items()}, d['parity'], d['cpu_baseline'])
P
tail -3 gpurun_out/r02_bench_h.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 700 gpurun_out/r02_bench_ref.json
