#!/bin/bash
# round-2 GPU call U (4 GPUs): weak and strong scaling bench lines at N = 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 4 --warmup 3 --no-extras --no-gpu-reference > gpurun_out/r02u_bench_n4_weak.json 2> gpurun_out/r02u_bench_n4_weak.err
echo "weak rc=$?"; head -c 400 gpurun_out/r02u_bench_n4_weak.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 4 --warmup 3 --scaling strong --no-extras --no-gpu-reference > gpurun_out/r02u_bench_n4_strong.json 2> gpurun_out/r02u_bench_n4_strong.err
echo "strong rc=$?"; head -c 400 gpurun_out/r02u_bench_n4_strong.json; echo
