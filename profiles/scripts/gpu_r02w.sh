#!/bin/bash
# round-2 GPU call W: 8 cached steps per graph replay (device-side token bookkeeping makes the step self-contained): parity tests + A/B bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_api_paths_gpu.py tests/test_model_gpu.py tests/test_golden_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02w_tests.log 2>&1; tail -4 gpurun_out/r02w_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02w_smoke.log
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02w_timeline --tag k8 > gpurun_out/r02w_tl.log 2>&1; tail -1 gpurun_out/r02w_tl.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-gpu-reference > gpurun_out/r02w_bench_k8.json 2> gpurun_out/r02w_bench_k8.err
echo "bench k8 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_k8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline_decode_step']['ms_per_decode_step'], d['gpu_launches'])"
AF3_DECODE_STEPS_PER_GRAPH=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-gpu-reference > gpurun_out/r02w_bench_k1.json 2> gpurun_out/r02w_bench_k1.err
echo "bench k1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_k1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline_decode_step']['ms_per_decode_step'], d['gpu_launches'])"
