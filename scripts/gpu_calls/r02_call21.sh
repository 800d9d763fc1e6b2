#!/bin/bash
# render tests after the near-plane change + probe for a GL implementation on the GPU box
mkdir -p gpurun_out
{
  echo "== GL probe"; ldconfig -p | grep -i -E "egl|libGL|gles|osmesa|glvnd|opengl|nvidia-gl|glcore" ; ls /usr/lib/x86_64-linux-gnu | grep -i -E "egl|libgl|glx|nvidia" | head -40
  echo "NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
  find / -xdev \( -name "libEGL*" -o -name "libGLX*" -o -name "libnvidia-egl*" -o -name "libnvidia-gl*" -o -name "libOSMesa*" \) 2>/dev/null | head -20
} > gpurun_out/r02_gl_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "render or tracker_with_cuda" 2>&1 | tail -15 > gpurun_out/r02_pytest_render.txt
cat gpurun_out/r02_pytest_render.txt
cat gpurun_out/r02_gl_probe.txt
