#!/bin/bash
# ncu --set full: the HBM-side kernels with the current K0 (half an image per 1024-thread CTA), and the trunk in latency mode (n = 1)
mkdir -p gpurun_out
export SE3TN_GRAPH=0
B="python bench.py --steps 2 --warmup 3 --no-alt --no-cpu-baseline --no-g21 --no-render"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_kernel|head_pooled_kernel" -s 6 -c 4 -o gpurun_out/r02_prof_aux_end $B > gpurun_out/r02_ncu_aux_end.log 2>&1
tail -2 gpurun_out/r02_ncu_aux_end.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_trunk_kernel" -s 3 -c 1 -o gpurun_out/r02_prof_trunk_n1 python scripts/trunk_units.py bf16x3 1 > gpurun_out/r02_ncu_trunk_n1.log 2>&1
tail -2 gpurun_out/r02_ncu_trunk_n1.log
ls -la gpurun_out/r02_prof_aux_end.ncu-rep gpurun_out/r02_prof_trunk_n1.ncu-rep
