#!/bin/bash
# round-2 GPU call O: 10-stage ring in the default few-token GEMM kernel (shared memory freed by the experiment split), K/V L2 prefetch on by default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "gemm or qkv or fused or decode_attention" > gpurun_out/r02o_tests_k.log 2>&1; tail -3 gpurun_out/r02o_tests_k.log
run() { tag=$1; shift; env "$@" timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02o_timeline --tag $tag > gpurun_out/r02o_tl.log 2>&1; tail -1 gpurun_out/r02o_tl.log; }
run s10_1 AF3_X=0
run s10_kv0 AF3_L2_PREFETCH_KV=0
run s10_gu4 AF3_L2_PREFETCH_GU=4
run s10_2 AF3_X=0
run s10_gu8 AF3_L2_PREFETCH_GU=8
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02o_tests_model.log 2>&1; tail -3 gpurun_out/r02o_tests_model.log
