#!/bin/bash
# round-2 GPU call V: greedy-loop bookkeeping on the device (af3_token_step inside the captured step): API-path + model parity tests, timeline, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_api_paths_gpu.py tests/test_model_gpu.py tests/test_golden_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02v_tests.log 2>&1; tail -4 gpurun_out/r02v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02v_smoke.log
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02v_timeline --tag tokstep > gpurun_out/r02v_tl.log 2>&1; tail -1 gpurun_out/r02v_tl.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02v_bench.json; echo
