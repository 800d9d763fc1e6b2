#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{ timeout 120 python scripts/trunk_units.py bf16x3 1; timeout 120 python scripts/trunk_units.py bf16x3 4; } > gpurun_out/r02_trunk_units2.txt 2>&1
cat gpurun_out/r02_trunk_units2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "latency or small_batches or on_track_end or config1" 2>&1 | tail -3
