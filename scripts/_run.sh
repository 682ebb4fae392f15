timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_gemm_gpu.py -m gpu -q 2>&1 | tail -3
DFU_TRACE=1 timeout 300 python scripts/trace_step.py mixed 1 > gpurun_out/trace_step_r02c.txt 2>&1; head -1 gpurun_out/trace_step_r02c.txt; tail -12 gpurun_out/trace_step_r02c.txt
DFU_WEIGHT_PREFETCH=0 DFU_TRACE=1 timeout 300 python scripts/trace_step.py mixed 1 > gpurun_out/trace_step_r02c_nopf.txt 2>&1; head -1 gpurun_out/trace_step_r02c_nopf.txt; tail -11 gpurun_out/trace_step_r02c_nopf.txt | head -4
python bench.py --steps 3 --warmup 3 --no-extra-configs > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -c 300 gpurun_out/r02j_bench.err
DFU_WEIGHT_PREFETCH=0 python bench.py --steps 3 --warmup 3 --no-extra-configs --no-cpu-baseline > gpurun_out/r02j_bench_nopf.json 2> gpurun_out/r02j_bench_nopf.err
python - <<'PY'
import json
for f in ('r02j_bench','r02j_bench_nopf'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['e2e']['value'], d['roofline']['unet_step_ms'], d['roofline']['frac'], d['roofline']['in_graph_ms_per_unet_step'])
PY
