#!/bin/bash
# round-2 GPU call K: no-split few-token GEMM (q/k/v, o projections): parity, sanitizers, isolated microbench, decode timeline A/B; gated xattn tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "gemm or qkv or fused" > gpurun_out/r02k_tests_gemm.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02k_tests_gemm.log; tail -4 gpurun_out/r02k_tests_gemm.log
timeout 600 python -m pytest tests/test_xattn_gpu.py -q -p no:cacheprovider > gpurun_out/r02k_tests_xattn.log 2>&1; tail -4 gpurun_out/r02k_tests_xattn.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02k_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "MISMATCH|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02k_sanitizer_$tool.log | tail -3
done
timeout 300 python profiles/microbench_fewtoken.py > gpurun_out/r02k_fewtoken.json 2> gpurun_out/r02k_fewtoken.err; echo "mb rc=$?"; grep -E "shape|kernel|us_per_launch|tail_us" gpurun_out/r02k_fewtoken.json | head -60
for i in 1 2; do
  timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02k_timeline --tag nosplit_$i > gpurun_out/r02k_tl_n$i.log 2>&1; tail -1 gpurun_out/r02k_tl_n$i.log
  AF3_SKINNY_MAXK=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02k_timeline --tag splitk_$i > gpurun_out/r02k_tl_s$i.log 2>&1; tail -1 gpurun_out/r02k_tl_s$i.log
done
AF3_SKINNY_MAXK=32768 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02k_timeline --tag nosplit_all > gpurun_out/r02k_tl_a.log 2>&1; tail -1 gpurun_out/r02k_tl_a.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_api_paths_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02k_tests_model.log 2>&1; tail -3 gpurun_out/r02k_tests_model.log
