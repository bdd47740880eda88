mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16_kernel -s 4 -c 1 -f -o gpurun_out/attn_x3_r2e python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_o1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16_kernel -s 4 -c 1 -f -o gpurun_out/attn_f16_r2e python tools/profile_step.py --precision f16 > gpurun_out/ncu_o2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16_kernel -s 4 -c 1 -f -o gpurun_out/attn_f16_c4_r2e python tools/profile_step.py --precision f16 --workload c4 > gpurun_out/ncu_o3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm_tf32_kernel -s 25 -c 1 -f -o gpurun_out/conv9_3x_r2e python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_o4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_3xf16_r2e.csv python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_o5.log 2>&1
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_o_ref.json 2> gpurun_out/bench_o_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_o.err | grep '^{"metric"' > gpurun_out/bench_o.json
python bench.py --steps 10 --warmup 3 --workload c4 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_o_c4.json
python bench.py --steps 10 --warmup 3 --workload c5 2>/dev/null | tail -1 > gpurun_out/bench_o_c5.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_o.log 2>&1
tail -4 gpurun_out/smoke_o.log; head -c 400 gpurun_out/bench_o_ref.json
