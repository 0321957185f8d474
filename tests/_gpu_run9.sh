timeout 1200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench7.json'))
print({k:d[k] for k in ['value','ms_per_step','audio_s_per_s','decode_tok_s','stage_ms','gpu_launches']})
print(d['roofline_decode_step'])
print({k:round(v,3) for k,v in d['decode_step_kernel_ms'].items()}, 'sum', round(sum(d['decode_step_kernel_ms'].values()),3))
for k in d['kernels']: print({a:(round(b,3) if isinstance(b,float) else b) for a,b in k.items() if a not in ('flops_per_launch','bytes_per_launch')})
PY
tail -5 gpurun_out/bench7.err | grep -v Warn
