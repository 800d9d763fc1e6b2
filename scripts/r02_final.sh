#!/bin/bash
# end-of-round validation (1 GPU): all GPU tests, smoke, the driver's two bench commands, the default bench, launch list + ncu --set full of the conv kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_final.txt
cat gpurun_out/r02_pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r02_smoke_final.txt
S=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike_ref.json 2> gpurun_out/r02_bench_driverlike_ref.err
echo "reference arm wall $(( $(date +%s) - S )) s"; tail -c 400 gpurun_out/r02_bench_driverlike_ref.json
S=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err
echo "our arm (20 steps) wall $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/r02_bench_driverlike.json
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
echo "default bench wall $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/r02_bench_final.json
B="python bench.py --steps 2 --warmup 3 --no-alt --no-cpu-baseline --no-g21 --no-render"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_final.csv $B > gpurun_out/r02_ncu_launches_final.log 2>&1
grep -c "kernel" gpurun_out/r02_launches_final.csv
export SE3TN_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_resident_kernel|conv_trunk_kernel" -s 27 -c 9 -o gpurun_out/r02_prof_conv_final $B > gpurun_out/r02_ncu_conv_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_kernel|head_pooled_kernel|render_kernel|render_project_kernel" -s 6 -c 4 -o gpurun_out/r02_prof_aux_final $B > gpurun_out/r02_ncu_aux_final.log 2>&1
ls -la gpurun_out/*final*
