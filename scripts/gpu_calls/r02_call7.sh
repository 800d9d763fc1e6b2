#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
for m in 0 1 5 13; do SE3TN_STEM_WS=$m timeout 120 python scripts/layer_times.py bf16x3; done
} > gpurun_out/r02_stem_ws_exp.txt 2>&1
cat gpurun_out/r02_stem_ws_exp.txt
