#!/bin/bash
# round-2 GPU call R: split-K tail with all partials (<= 5 splits) and the residual requested in one round trip
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "gemm or qkv or fused" > gpurun_out/r02r_tests_k.log 2>&1; tail -3 gpurun_out/r02r_tests_k.log
for i in 1 2; do timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02r_timeline --tag onert_$i > gpurun_out/r02r_tl.log 2>&1; tail -1 gpurun_out/r02r_tl.log; done
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -k "greedy or pdl or graph" > gpurun_out/r02r_tests_model.log 2>&1; tail -3 gpurun_out/r02r_tests_model.log
