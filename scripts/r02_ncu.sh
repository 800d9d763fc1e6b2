#!/bin/bash
# round 2 evidence run (1 GPU): launch list, ncu --set full of the conv kernels (batch 64 bf16x3, batch 256 bf16) and of the HBM-side kernels, racecheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SE3TN_GRAPH=0
B="python bench.py --steps 2 --warmup 3 --no-alt --no-cpu-baseline --no-g21 --no-render"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_resident_kernel|conv_trunk_kernel|conv_stem_ws_kernel" -s 27 -c 9 -o gpurun_out/r02_prof_conv $B > gpurun_out/r02_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_kernel|head_pooled_kernel" -s 6 -c 4 -o gpurun_out/r02_prof_aux2 $B > gpurun_out/r02_ncu_aux2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_resident_kernel|conv_trunk_kernel|conv_stem_ws_kernel" -s 27 -c 9 -o gpurun_out/r02_prof_conv_b256 $B --batch 256 --precision bf16 > gpurun_out/r02_ncu_conv_b256.log 2>&1
unset SE3TN_GRAPH
timeout 600 python bench.py --batch 256 --precision bf16 --steps 20 --warmup 5 --no-alt --no-g21 --no-render --no-cpu-baseline > gpurun_out/r02_bench_b256_bf16.json 2> gpurun_out/r02_bench_b256_bf16.err
tail -c 600 gpurun_out/r02_bench_b256_bf16.json
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/r02_racecheck2.txt 2>&1
tail -4 gpurun_out/r02_racecheck2.txt
ls -la gpurun_out/*.ncu-rep
