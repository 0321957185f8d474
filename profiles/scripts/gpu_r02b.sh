#!/bin/bash
# round-2 GPU call B: suite again (fixed tests, attention v2, split-K tail changes, packed rmsnorm), split-K sweep, attention A/B,
# decode timeline, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -4 gpurun_out/r02b_tests.log
timeout 300 python profiles/microbench_attention.py > gpurun_out/r02b_attn.json 2> gpurun_out/r02b_attn.err; echo "attn rc=$?"; cat gpurun_out/r02b_attn.json | head -40
timeout 400 python profiles/microbench_splitk.py > gpurun_out/r02b_splitk.json 2> gpurun_out/r02b_splitk.err; echo "splitk rc=$?"
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02b_timeline --tag tail_opt > gpurun_out/r02b_timeline.log 2>&1; tail -1 gpurun_out/r02b_timeline.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; head -c 400 gpurun_out/r02b_bench.json
