#!/bin/bash
# round-2 GPU call I: packed-math fused RMSNorm transform: parity + alternating A/B timelines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -x -k "fused or pdl or greedy or splitk or swap or swiglu" > gpurun_out/r02i_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02i_tests.log; tail -3 gpurun_out/r02i_tests.log
for i in 1 2; do
  timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02i_timeline --tag fused_$i > gpurun_out/r02i_tl_f$i.log 2>&1; tail -1 gpurun_out/r02i_tl_f$i.log
  AF3_FUSE_NORM=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02i_timeline --tag unfused_$i > gpurun_out/r02i_tl_u$i.log 2>&1; tail -1 gpurun_out/r02i_tl_u$i.log
done
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "attention" > gpurun_out/r02i_tests_attn.log 2>&1; tail -3 gpurun_out/r02i_tests_attn.log
timeout 300 python profiles/microbench_attention.py > gpurun_out/r02i_attn.json 2> gpurun_out/r02i_attn.err; echo "attn rc=$?"; grep -E "tflops|speedup" gpurun_out/r02i_attn.json
