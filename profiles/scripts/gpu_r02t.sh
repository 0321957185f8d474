#!/bin/bash
# round-2 GPU call T (2 GPUs): weak and strong scaling bench lines at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-extras --no-gpu-reference > gpurun_out/r02t_bench_n2_weak.json 2> gpurun_out/r02t_bench_n2_weak.err
echo "weak rc=$?"; head -c 400 gpurun_out/r02t_bench_n2_weak.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --scaling strong --no-extras --no-gpu-reference > gpurun_out/r02t_bench_n2_strong.json 2> gpurun_out/r02t_bench_n2_strong.err
echo "strong rc=$?"; head -c 400 gpurun_out/r02t_bench_n2_strong.json; echo
