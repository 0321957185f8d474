#!/bin/bash
# round-2 GPU call S: split-K reduction through bulk copies of the partial tiles into the idle shared-memory ring (one round trip)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "gemm or qkv or fused" > gpurun_out/r02s_tests_k.log 2>&1; tail -3 gpurun_out/r02s_tests_k.log
for i in 1 2; do timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02s_timeline --tag bulk_$i > gpurun_out/r02s_tl.log 2>&1; tail -1 gpurun_out/r02s_tl.log; done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02s_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "MISMATCH|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02s_sanitizer_$tool.log | tail -3
done
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -k "greedy or pdl or graph" > gpurun_out/r02s_tests_model.log 2>&1; tail -3 gpurun_out/r02s_tests_model.log
