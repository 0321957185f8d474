#!/bin/bash
# round-2 GPU call G: warp-granular fused RMSNorm, [gate; up] views, cluster reduce opt-in; alternating A/B timelines on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02g_tests.log; tail -5 gpurun_out/r02g_tests.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py > gpurun_out/r02g_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/r02g_sanitizer_$tool.log
done
for i in 1 2; do
  timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02g_timeline --tag fused_$i > gpurun_out/r02g_tl_f$i.log 2>&1; tail -1 gpurun_out/r02g_tl_f$i.log
  AF3_FUSE_NORM=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02g_timeline --tag unfused_$i > gpurun_out/r02g_tl_u$i.log 2>&1; tail -1 gpurun_out/r02g_tl_u$i.log
done
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02g_bench.json
