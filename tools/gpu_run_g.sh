mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "attention or golden or oracle or edge or inference or random_shapes or graph" 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 > gpurun_out/pytest_g.log
FS2_ATT_TRACE=1 timeout 300 python tools/attn_probe.py 2> gpurun_out/attn_trace_g.log | tail -3 > gpurun_out/attn_probe_g.log
timeout 300 python tools/attn_probe.py 32 2000 384 2 2>/dev/null | tail -3 >> gpurun_out/attn_probe_g.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
timeout 600 python bench.py --steps 20 --warmup 5 --e2e-first 1 --modes '' > gpurun_out/bench_g_e2efirst.json 2> /dev/null
tail -8 gpurun_out/pytest_g.log; cat gpurun_out/attn_probe_g.log
