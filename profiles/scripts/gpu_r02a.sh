#!/bin/bash
# round-2 GPU call A: full GPU test suite (incl. the new full-depth parity + API-path tests), decode-step timelines for the
# pipeline-depth variants, default bench (with the HF-on-GPU reference and the extras)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r02a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
run_tl() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02a_timeline --tag "$tag" > gpurun_out/r02a_timeline_$tag.log 2>&1
  tail -1 gpurun_out/r02a_timeline_$tag.log
}
run_tl base AF3_X=0
run_tl s4_3 AF3_SWAP_STAGES=4 AF3_SWAP_STAGES2=3
run_tl s3_3_da2 AF3_SWAP_STAGES=3 AF3_SWAP_STAGES2=3 AF3_DA_STAGES=2
run_tl s4_3_da2 AF3_SWAP_STAGES=4 AF3_SWAP_STAGES2=3 AF3_DA_STAGES=2
run_tl pdl0 AF3_PDL=0
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/r02a_bench.json
