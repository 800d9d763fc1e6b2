#!/bin/bash
# n = 1 latency breakdown + ncu capture of the rasteriser after the near-plane change
mkdir -p gpurun_out
timeout 600 python scripts/latency_breakdown.py > gpurun_out/r02_latency_breakdown.txt 2>&1
cat gpurun_out/r02_latency_breakdown.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_kernel|render_project_kernel" -s 4 -c 2 -o gpurun_out/r02_prof_render python scripts/render_once.py > gpurun_out/r02_ncu_render.log 2>&1
tail -3 gpurun_out/r02_ncu_render.log
