#!/bin/bash
# two-GPU evidence run (gpurun --gpus 2): NCCL C-ABI test, sharded == single-GPU identity test, 2-GPU bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multigpu.py -m gpu -q -s -v > gpurun_out/r02_pytest_2gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_2gpu.txt
tail -15 gpurun_out/r02_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 1500 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err
