#!/bin/bash
# sanitizers over the small all-kernels script (latency mode with owned blocks, pyrender-style render, track_host)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/r02_racecheck3.txt 2>&1
tail -4 gpurun_out/r02_racecheck3.txt
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_memcheck3.txt 2>&1
tail -4 gpurun_out/r02_memcheck3.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "track_host" 2>&1 | tail -3
