#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_i.json').read().strip().splitlines()[-1])
print('bench i:', d['value'], d['ms_per_step'], d.get('graph_launches_per_step'), d['single_track']['ms_per_frame'], d['e2e']['value'], d['weight_sets_21']['ratio_vs_1_set'], {k:(v['value'], v['ms_per_step']) for k,v in d['alt_precisions'].items()})
P
tail -3 gpurun_out/r02_bench_i.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "on_track_end or tracker_with_cuda" 2>&1 | tail -3
