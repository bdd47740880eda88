mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_b.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_b.log 2>&1
python bench.py --steps 40 --warmup 3 --modes '' > gpurun_out/bench_b_3x.json 2> gpurun_out/bench_b_3x.err
python bench.py --steps 20 --warmup 3 --precision f16 --modes '' > gpurun_out/bench_b_f16.json 2> gpurun_out/bench_b_f16.err
python bench.py --steps 10 --warmup 3 --workload c4 --modes f16 > gpurun_out/bench_b_c4.json 2> gpurun_out/bench_b_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_3xf16_r2.csv python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16_kernel -s 4 -c 1 -f -o gpurun_out/attn_x3_r2 python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm_tf32_kernel -s 25 -c 1 -f -o gpurun_out/conv9_3x_r2 python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16x2_kernel -c 1 -f -o gpurun_out/attn_f16x2_r2 python tools/profile_step.py --precision f16 > gpurun_out/ncu_x.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ln_tf32_kernel -c 1 -f -o gpurun_out/gemm_ln_f16_r2 python tools/profile_step.py --precision f16 > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/pytest_b.log; cat gpurun_out/smoke_b.log | tail -5
