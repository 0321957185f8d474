#!/bin/bash
# round-2 GPU call N: selective L2 prefetch before the dependency resolves (K/V of decode attention only; gate/up weights only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02n_timeline --tag $tag > gpurun_out/r02n_tl.log 2>&1; tail -1 gpurun_out/r02n_tl.log; }
run base_1 AF3_X=0
run kv_1 AF3_L2_PREFETCH_KV=1
run kv_gu4_1 AF3_L2_PREFETCH_KV=1 AF3_L2_PREFETCH_GU=4
run gu4_1 AF3_L2_PREFETCH_GU=4
run base_2 AF3_X=0
run kv_2 AF3_L2_PREFETCH_KV=1
run kv_gu2_1 AF3_L2_PREFETCH_KV=1 AF3_L2_PREFETCH_GU=2
