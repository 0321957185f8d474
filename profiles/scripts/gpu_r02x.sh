#!/bin/bash
# round-2 GPU call X (8 GPUs): weak-scaling bench line at N = 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 --no-extras --no-gpu-reference > gpurun_out/r02x_bench_n8_weak.json 2> gpurun_out/r02x_bench_n8_weak.err
echo "weak rc=$?"; head -c 400 gpurun_out/r02x_bench_n8_weak.json; echo
