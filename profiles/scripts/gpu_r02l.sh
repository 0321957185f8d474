#!/bin/bash
# round-2 GPU call L: lean default few-token GEMM kernels (experiments compiled out): full GPU tests, decode timelines, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/r02l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02l_tests.log; tail -4 gpurun_out/r02l_tests.log
for i in 1 2; do
  timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02l_timeline --tag lean_$i > gpurun_out/r02l_tl_$i.log 2>&1; tail -1 gpurun_out/r02l_tl_$i.log
done
AF3_FUSE_NORM=1 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02l_timeline --tag fused_norm > gpurun_out/r02l_tl_f.log 2>&1; tail -1 gpurun_out/r02l_tl_f.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
echo "bench rc=$?"; head -c 400 gpurun_out/r02l_bench.json; echo
