mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ln_cluster -s 8 -c 2 -f -o gpurun_out/gemm_ln_cl_x3 python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_s1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ln_cluster -s 8 -c 2 -f -o gpurun_out/gemm_ln_cl_f16 python tools/profile_step.py --precision f16 > gpurun_out/ncu_s2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_3xf16_r2f.csv python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_s3.log 2>&1
tail -2 gpurun_out/ncu_s1.log gpurun_out/ncu_s2.log gpurun_out/ncu_s3.log
