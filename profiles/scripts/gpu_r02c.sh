#!/bin/bash
# round-2 GPU call C: compute-sanitizer over every kernel at small shapes; ncu --set full (source-level) of the v2 attention kernels
# and of the decode attention kernel; the fixed EOS test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02c_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02c_sanitizer_$tool.log
done
AF3_MB_CASE=prefill_d128_t780 AF3_MB_IMPL=v2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 3 -c 1 -f -o gpurun_out/r02c_attn2_d128 python profiles/microbench_attention.py > gpurun_out/r02c_ncu_attn128.log 2>&1; echo "ncu128 rc=$?"
AF3_MB_CASE=encoder AF3_MB_IMPL=v2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 3 -c 1 -f -o gpurun_out/r02c_attn2_d64 python profiles/microbench_attention.py > gpurun_out/r02c_ncu_attn64.log 2>&1; echo "ncu64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_attn_kernel -s 2 -c 1 -f -o gpurun_out/r02c_dattn python profiles/decode_timeline.py --layers 4 --new-tokens 6 --out gpurun_out/r02c_ncu_tl --tag ncu > gpurun_out/r02c_ncu_dattn.log 2>&1; echo "ncu dattn rc=$?"
timeout 600 python -m pytest tests/test_api_paths_gpu.py -q -p no:cacheprovider -k "eos or rejects" > gpurun_out/r02c_tests.log 2>&1; tail -3 gpurun_out/r02c_tests.log
ls -la gpurun_out/*.ncu-rep
