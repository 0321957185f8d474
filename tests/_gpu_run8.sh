timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6.json'))
print({k:d[k] for k in ['value','ms_per_step','audio_s_per_s','decode_tok_s','stage_ms','gpu_launches']})
print(d['roofline_decode_step'])
PY
AF3_PDL=0 timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_nopdl.json 2>> gpurun_out/bench6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6_nopdl.json'))
print("NO PDL", {k:d[k] for k in ['value','ms_per_step','decode_tok_s','stage_ms']})
PY
tail -5 gpurun_out/bench6.err | grep -v Warn
