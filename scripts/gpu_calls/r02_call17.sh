#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/graph_alt_debug.py > gpurun_out/r02_graph_alt_debug.txt 2>&1; cat gpurun_out/r02_graph_alt_debug.txt
