#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{ timeout 120 python scripts/trunk_units.py bf16x3 1; timeout 120 python scripts/trunk_units.py bf16x3 8; } > gpurun_out/r02_trunk_units.txt 2>&1
cat gpurun_out/r02_trunk_units.txt
