timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
for shp in "64 64 320 8" "64 64 960 8" "32 32 640 8" "128 128 512 2"; do DFU_TRACE=1 timeout 120 python scripts/bench_gn.py $shp 2>&1 | grep -E "GroupNorm|auto"; done
python bench.py --steps 3 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 300 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['unet_step_ms'], d['roofline']['frac'], d['roofline']['in_graph_ms_per_unet_step'])
for k,v in d['configs'].items(): print(k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if not isinstance(vv,(dict,str))}, v.get('in_graph_ms_per_unet_step'), (v.get('roofline_gemm') or {}).get('frac'))
PY
