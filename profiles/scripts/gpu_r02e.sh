#!/bin/bash
# round-2 GPU call E: cluster / DSMEM split-K reduction + fused RMSNorm v2 (token-major partials) + attention v2 without chunk unrolling
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_golden_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log; tail -4 gpurun_out/r02e_tests.log
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02e_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/r02e_sanitizer_$tool.log
done
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02e_timeline --tag cluster_fused > gpurun_out/r02e_tl1.log 2>&1; tail -1 gpurun_out/r02e_tl1.log
AF3_CLUSTER_REDUCE=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02e_timeline --tag global_fused > gpurun_out/r02e_tl2.log 2>&1; tail -1 gpurun_out/r02e_tl2.log
AF3_CLUSTER_REDUCE=0 AF3_FUSE_NORM=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02e_timeline --tag round1_chain > gpurun_out/r02e_tl3.log 2>&1; tail -1 gpurun_out/r02e_tl3.log
timeout 500 python profiles/microbench_splitk.py > gpurun_out/r02e_splitk.json 2> gpurun_out/r02e_splitk.err; echo "splitk rc=$?"
timeout 300 python profiles/microbench_attention.py > gpurun_out/r02e_attn.json 2> gpurun_out/r02e_attn.err; echo "attn rc=$?"; grep -E "tflops|speedup" gpurun_out/r02e_attn.json
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02e_bench.json
