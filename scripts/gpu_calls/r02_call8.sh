#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/stem_ws_check.py > gpurun_out/r02_stem_ws2.txt 2>&1; cat gpurun_out/r02_stem_ws2.txt
SE3TN_STEM_WS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or small_batches or batch64 or raw_regime or batch256 or mixed_weight or on_track_end" > gpurun_out/r02_pytest_ws.txt 2>&1; tail -5 gpurun_out/r02_pytest_ws.txt
for m in 0 1; do SE3TN_STEM_WS=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-g21 --no-render > gpurun_out/r02_bench_e_$m.json 2> gpurun_out/r02_bench_e_$m.err; done
python - <<'P'
import json
for m in (0,1):
    d=json.loads(open('gpurun_out/r02_bench_e_%d.json'%m).read().strip().splitlines()[-1])
    print('stem_ws', m, d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms']['conv'][:3], d['roofline']['per_kernel_ms']['preprocess'])
P
