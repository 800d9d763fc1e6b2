#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ycb or headless" 2>&1 | tail -15 > gpurun_out/r02_pytest_call36.txt
cat gpurun_out/r02_pytest_call36.txt
