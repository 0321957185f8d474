#!/bin/bash
# round-2 GPU call M: decode attention with two head groups; L2 prefetch of weights / K,V before the dependency resolves (sweep)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "decode_attention or gemm or qkv" > gpurun_out/r02m_tests_k.log 2>&1; tail -3 gpurun_out/r02m_tests_k.log
timeout 300 python profiles/microbench_decode_attn.py > gpurun_out/r02m_dattn.json 2> gpurun_out/r02m_dattn.err; echo "dattn rc=$?"; cat gpurun_out/r02m_dattn.json | head -30
for pf in 0 8 16 32 0 16; do
  AF3_L2_PREFETCH=$pf timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02m_timeline --tag pf${pf}_$RANDOM > gpurun_out/r02m_tl.log 2>&1; tail -1 gpurun_out/r02m_tl.log
done
AF3_L2_PREFETCH=16 timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02m_tests_model_pf.log 2>&1; tail -3 gpurun_out/r02m_tests_model_pf.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_api_paths_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02m_tests_model.log 2>&1; tail -3 gpurun_out/r02m_tests_model.log
