set -x
NCU="ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_unet_vae_mixed_r02.csv python scripts/profile_step.py mixed 1 vae > gpurun_out/ncu1.log 2>&1
$NCU --profile-from-start off -k regex:'gn_|layernorm|cast_kernel|conv_small|splitk|attn_merge|softmax|transpose' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv --log-file gpurun_out/norm_kernels_r02.csv python scripts/profile_step.py mixed 1 vae > gpurun_out/ncu2.log 2>&1
$NCU --profile-from-start off -k regex:'gemm' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/gemm_dram_r02.csv python scripts/profile_step.py mixed 1 > gpurun_out/ncu3.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_conv_b1 python scripts/profile_gemm_one.py > gpurun_out/ncu4.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_conv_b8 python scripts/profile_gemm_one.py --batch 8 --tune 160,1,3,1 > gpurun_out/ncu5.log 2>&1
$NCU --set full --import-source on -k regex:gemm2_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm2_vae python scripts/profile_gemm_one.py --hw 256 --cin 256 --cout 256 --prec 2 --tune 256,1,0,2 > gpurun_out/ncu6.log 2>&1
$NCU --set full --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/prof_attn_l0 python scripts/profile_gemm_one.py --attn > gpurun_out/ncu7.log 2>&1
$NCU --set full -k regex:'gn_cluster|layernorm|cast_kernel|conv_small_out' -s 10 -c 5 -f -o gpurun_out/prof_norm_b1 python scripts/profile_norm_one.py 1 > gpurun_out/ncu8.log 2>&1
$NCU --set full -k regex:'gn_cluster|layernorm|cast_kernel|conv_small_out' -s 10 -c 5 -f -o gpurun_out/prof_norm_b8 python scripts/profile_norm_one.py 8 > gpurun_out/ncu9.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*_r02.csv
tail -n 2 gpurun_out/ncu*.log
