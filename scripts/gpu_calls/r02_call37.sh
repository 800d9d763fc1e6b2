#!/bin/bash
# preprocess with half an image per 1024-thread CTA at n >= 64: all GPU tests + the driver's bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_call37.txt
cat gpurun_out/r02_pytest_call37.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_call37.json 2> gpurun_out/r02_bench_call37.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_call37.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['roofline']['per_kernel_ms']['preprocess'], d['single_track']['ms_per_frame'], d['weight_sets_21']['ratio_vs_1_set'])
PY
timeout 300 python bench.py --batch 256 --precision bf16 --steps 20 --warmup 5 --no-alt --no-g21 --no-render --no-cpu-baseline > gpurun_out/r02_bench_b256_call37.json 2>/dev/null
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_b256_call37.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('b256 bf16', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_kernel_ms']['preprocess'], d['e2e']['value'])
PY
