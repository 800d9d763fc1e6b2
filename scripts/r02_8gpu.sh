#!/bin/bash
# N-GPU bench line exactly as the driver launches it (N = number of visible GPUs): ours, then the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "rc=$? wall $(( $(date +%s) - S )) s"
tail -c 1800 gpurun_out/r02_bench_${N}gpu.json; tail -5 gpurun_out/r02_bench_${N}gpu.err
S=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_bench_${N}gpu_ref.json 2> gpurun_out/r02_bench_${N}gpu_ref.err
echo "reference arm rc=$? wall $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/r02_bench_${N}gpu_ref.json
