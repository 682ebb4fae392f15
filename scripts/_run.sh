timeout 900 python -m pytest tests/test_norm_misc_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 300 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['unet_step_ms'], d['roofline']['frac'], d['roofline']['in_graph_ms_per_unet_step'])
for k,v in d['configs'].items(): print(k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if not isinstance(vv,(dict,str))}, v.get('in_graph_ms_per_unet_step'), (v.get('roofline_gemm') or {}).get('frac'))
PY
