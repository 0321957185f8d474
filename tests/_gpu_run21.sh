timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "decode_attention" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "generate or pdl" 2>&1 | tail -3
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
NCU="ncu --profile-from-start off --clock-control none"
AF3_NCU_GRAPH=1 AF3_NCU_NEW_TOKENS=6 AF3_PDL=0 timeout 900 $NCU --graph-profiling node --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/launches_r01g_graph.csv python profiles/ncu_driver.py > gpurun_out/ncu_list2.log 2>&1; echo "graph list rc=$?"
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench18.json'))
print({k:d[k] for k in ['value','ms_per_step','decode_tok_s','stage_ms','gpu_launches']}, d['roofline_decode_step']['ms_per_decode_step'])
PY
