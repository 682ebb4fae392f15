set -x
mkdir -p gpurun_out
TUNE_OUT=unet_b1 timeout 700 python scripts/tune_insitu.py 1 512 unet > gpurun_out/tune_insitu_unet_b1_r02b.txt 2>&1; tail -n 2 gpurun_out/tune_insitu_unet_b1_r02b.txt
if [ -s gpurun_out/tuning_b200_unet_b1.json ]; then cp gpurun_out/tuning_b200_unet_b1.json diffute_b200/tuning_b200.json; fi
TUNE_OUT=unet_b8 timeout 900 python scripts/tune_insitu.py 8 512 unet > gpurun_out/tune_insitu_unet_b8_r02b.txt 2>&1; tail -n 2 gpurun_out/tune_insitu_unet_b8_r02b.txt
