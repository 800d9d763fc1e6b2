#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
python scripts/trace_timeline.py bf16x3 64
python scripts/trace_timeline.py bf16x3 1
python scripts/trace_timeline.py bf16 64
SE3TN_PDL=0 python scripts/trace_timeline.py bf16x3 64
} > gpurun_out/r02_trace.txt 2>&1
cat gpurun_out/r02_trace.txt
