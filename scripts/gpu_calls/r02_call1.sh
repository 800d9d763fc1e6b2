#!/bin/bash
# round 2, GPU call 1: where does the time go in the N=64 kernels (epilogue vs MMA), ncu evidence for the HBM-side kernels, racecheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
for p in bf16x3 bf16; do
  python scripts/layer_times.py $p
  SE3TN_DEBUG_SKIP=4 python scripts/layer_times.py $p
  SE3TN_DEBUG_SKIP=8 python scripts/layer_times.py $p
  SE3TN_DEBUG_SKIP=2 python scripts/layer_times.py $p
done
SE3TN_PDL=0 python scripts/layer_times.py bf16x3
} > gpurun_out/r02_layer_times.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_kernel|head_pooled_kernel|pose_update_kernel" -s 9 -c 6 \
  -o gpurun_out/r02_prof_aux python bench.py --steps 2 --warmup 3 --no-alt --no-cpu-baseline --no-render > gpurun_out/r02_ncu_aux.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/r02_racecheck.txt 2>&1
tail -5 gpurun_out/r02_racecheck.txt
cat gpurun_out/r02_layer_times.txt
