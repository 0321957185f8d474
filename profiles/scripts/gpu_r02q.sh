#!/bin/bash
# round-2 GPU call Q: validation of HEAD + the evidence set.  ncu reports are summarised ON THE BOX and deleted (gpurun_out/ is capped at 64 MiB).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r02q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02q_tests.log; tail -4 gpurun_out/r02q_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02q_smoke.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02q_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -1 gpurun_out/r02q_sanitizer_$tool.log
done
AF3_NCU_NEW_TOKENS=4 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02q_launches.csv python profiles/ncu_driver.py > gpurun_out/r02q_ncu_launches.log 2>&1; echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, note, command...
  name=$1; kre=$2; skip=$3; note=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/r02q_$name "$@" > gpurun_out/r02q_ncu_$name.log 2>&1
  rc=$?
  if [ -f gpurun_out/r02q_$name.ncu-rep ]; then
    python profiles/ncu_source_summary.py gpurun_out/r02q_$name.ncu-rep gpurun_out/r02q_ncu_$name.md "$note" > /dev/null 2>&1
    rm -f gpurun_out/r02q_$name.ncu-rep
  fi
  echo "ncu $name rc=$rc"
}
cap gemm_prefill_gateup gemm_kernel 2 "prefill gate/up 24960 x 2*18944 x 3584, SwiGLU" python profiles/ncu_targets.py prefill_gateup
cap gemm_prefill_down gemm_kernel 2 "prefill down 24960 x 3584 x 18944 + residual" python profiles/ncu_targets.py prefill_down
cap gemm_decode_gateup gemm_kernel 2 "decode gate/up, 32 tokens" python profiles/ncu_targets.py decode_gateup
cap gemm_decode_down gemm_kernel 2 "decode down, 32 tokens, split-K x5 + residual" python profiles/ncu_targets.py decode_down
cap gemm_decode_qkv gemm_kernel 2 "decode q/k/v, 32 tokens, split-K x4 + bias" python profiles/ncu_targets.py decode_qkv
cap gemm_decode_o gemm_kernel 2 "decode o, 32 tokens, split-K x5 + residual" python profiles/ncu_targets.py decode_o
AF3_MB_CASE=prefill_d128_t780 AF3_MB_IMPL=v2 cap attn2_d128 attention2_kernel 3 "prefill attention D=128, T=780, causal GQA" python profiles/microbench_attention.py
AF3_MB_CASE=encoder AF3_MB_IMPL=v2 cap attn2_d64 attention2_kernel 3 "encoder attention D=64, T=1500" python profiles/microbench_attention.py
AF3_MB_EAGER=1 cap decode_attn decode_attn_kernel 30 "decode attention, 32 x 4 KV heads, ctx 780, two head groups" python profiles/microbench_decode_attn.py
timeout 300 python profiles/microbench_attention.py > gpurun_out/r02q_attn.json 2> gpurun_out/r02q_attn.err
timeout 300 python profiles/microbench_decode_attn.py > gpurun_out/r02q_dattn.json 2> gpurun_out/r02q_dattn.err
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02q_timeline --tag head > gpurun_out/r02q_tl.log 2>&1; tail -1 gpurun_out/r02q_tl.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02q_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02q_bench_reference.json 2> gpurun_out/r02q_bench_reference.err; echo "ref rc=$?"
du -sh gpurun_out
