#!/bin/bash
# round-2 GPU call J: validation of HEAD + the evidence set (tests, sanitizers, ncu launch list + --set full captures, bench + reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r02j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02j_tests.log; tail -4 gpurun_out/r02j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02j_smoke.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02j_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -1 gpurun_out/r02j_sanitizer_$tool.log
done
AF3_NCU_NEW_TOKENS=4 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02j_launches.csv python profiles/ncu_driver.py > gpurun_out/r02j_ncu_launches.log 2>&1; echo "launch list rc=$?"
for c in prefill_gateup prefill_down decode_gateup decode_down decode_qkv; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/r02j_gemm_$c python profiles/ncu_targets.py $c > gpurun_out/r02j_ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
AF3_MB_CASE=prefill_d128_t780 AF3_MB_IMPL=v2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 3 -c 1 -f -o gpurun_out/r02j_attn2_d128 python profiles/microbench_attention.py > gpurun_out/r02j_ncu_attn128.log 2>&1; echo "ncu attn128 rc=$?"
AF3_MB_CASE=encoder AF3_MB_IMPL=v2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 3 -c 1 -f -o gpurun_out/r02j_attn2_d64 python profiles/microbench_attention.py > gpurun_out/r02j_ncu_attn64.log 2>&1; echo "ncu attn64 rc=$?"
timeout 300 python profiles/microbench_attention.py > gpurun_out/r02j_attn.json 2> gpurun_out/r02j_attn.err
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02j_timeline --tag head > gpurun_out/r02j_tl.log 2>&1; tail -1 gpurun_out/r02j_tl.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02j_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err; echo "ref rc=$?"
ls -la gpurun_out/*.ncu-rep | head -12
