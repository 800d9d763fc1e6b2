#!/bin/bash
# pyrender-style mode of the rasteriser: render tests (both modes) + render timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "render or tracker" 2>&1 | tail -25 > gpurun_out/r02_pytest_render2.txt
cat gpurun_out/r02_pytest_render2.txt
timeout 300 python scripts/render_once.py 2>&1 | tail -3
