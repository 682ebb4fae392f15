timeout 200 python scripts/exp_vae.py 8 2>&1 | grep -v Warn
echo "--- DFU_GN_SMEM=0"; DFU_GN_SMEM=0 timeout 200 python scripts/exp_vae.py 8 2>&1 | grep -v Warn
echo "--- DFU_GEMM_AUTO_PAIR=0"; DFU_GEMM_AUTO_PAIR=0 timeout 200 python scripts/exp_vae.py 8 2>&1 | grep -v Warn
echo "--- B=1"; timeout 200 python scripts/exp_vae.py 1 2>&1 | grep -v Warn
echo "--- GN B=1 reg vs smem"
for shp in "64 64 320 1" "32 32 640 1" "16 16 1280 1" "64 64 960 1" "32 32 1920 1"; do DFU_TRACE=1 timeout 120 python scripts/bench_gn.py $shp 2>&1 | grep -E "GroupNorm|auto"; DFU_GN_SMEM=2 DFU_TRACE=1 timeout 120 python scripts/bench_gn.py $shp 2>&1 | grep -E "auto"; done
