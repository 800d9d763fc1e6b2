#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/single_track_ab.py > gpurun_out/r02_single_track_ab.txt 2>&1; head -60 gpurun_out/r02_single_track_ab.txt | cut -c1-200
