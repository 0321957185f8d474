#!/bin/bash
# round-2 GPU call D: fused RMSNorm across the decode GEMMs (tests, timeline A/B), synccheck again after the o_full change
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_golden_gpu.py -q -p no:cacheprovider -x > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02d_tests.log; tail -4 gpurun_out/r02d_tests.log
timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02d_timeline --tag fused_norm > gpurun_out/r02d_tl1.log 2>&1; tail -1 gpurun_out/r02d_tl1.log
AF3_FUSE_NORM=0 timeout 400 python profiles/decode_timeline.py --out gpurun_out/r02d_timeline --tag unfused > gpurun_out/r02d_tl0.log 2>&1; tail -1 gpurun_out/r02d_tl0.log
timeout 900 compute-sanitizer --tool synccheck python profiles/sanitize_kernels.py > gpurun_out/r02d_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/r02d_sanitizer_synccheck.log
timeout 600 python -m pytest tests/test_api_paths_gpu.py tests/test_parity_full_gpu.py -q -p no:cacheprovider > gpurun_out/r02d_tests2.log 2>&1; tail -3 gpurun_out/r02d_tests2.log
