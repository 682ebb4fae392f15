for shp in "64 64 320 8" "64 64 960 8" "32 32 640 8" "32 32 1920 8" "16 16 1280 8" "64 64 320 2" "32 32 640 2"; do DFU_GN_CLUSTER=0 DFU_TRACE=1 timeout 120 python scripts/bench_gn.py $shp 2>&1 | grep -E "GroupNorm|auto"; DFU_TRACE=1 timeout 120 python scripts/bench_gn.py $shp 2>&1 | grep -E "auto"; done
DFU_TRACE=1 timeout 300 python scripts/trace_step.py mixed 8 > gpurun_out/trace_step_b8_r02b.txt 2>&1
tail -12 gpurun_out/trace_step_b8_r02b.txt
