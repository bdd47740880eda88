mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py tests/test_vocoder.py tests/test_data.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -80 > gpurun_out/pytest_d.log
FS2_ATT_X2=0 python bench.py --steps 20 --warmup 3 --precision f16 --modes '' > gpurun_out/bench_d_f16_x1.json 2> gpurun_out/bench_d_f16_x1.err
FS2_ATT_X2=1 python bench.py --steps 20 --warmup 3 --precision f16 --modes '' > gpurun_out/bench_d_f16_x2.json 2> gpurun_out/bench_d_f16_x2.err
FS2_ATT_X2=0 python bench.py --steps 10 --warmup 3 --precision f16 --workload c4 --modes '' > gpurun_out/bench_d_c4_f16_x1.json 2> /dev/null
FS2_ATT_X2=1 python bench.py --steps 10 --warmup 3 --precision f16 --workload c4 --modes '' > gpurun_out/bench_d_c4_f16_x2.json 2> /dev/null
tail -30 gpurun_out/pytest_d.log
