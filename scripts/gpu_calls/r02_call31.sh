#!/bin/bash
# racecheck again after making the scheduler-ring read single-threaded; latency + host-call tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/r02_racecheck4.txt 2>&1
tail -4 gpurun_out/r02_racecheck4.txt
grep -n "in conv_\|in aux_\|in render\|in depth_\|in metrics" gpurun_out/r02_racecheck4.txt | sed 's/.* in //' | sort | uniq -c | head
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "latency or track_host or config1 or batch64" 2>&1 | tail -3
