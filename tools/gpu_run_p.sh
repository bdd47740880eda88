mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm" 2>&1 | tail -15 > gpurun_out/pytest_p_ln.log
cat gpurun_out/pytest_p_ln.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_p.log
cat gpurun_out/pytest_p.log
for cl in 1 0; do
  FS2_LN_CLUSTER=$cl timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_p_cl$cl.json
  FS2_LN_CLUSTER=$cl timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --precision f16 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_p_f16_cl$cl.json
done
python - <<'PY'
import json
for f in ["bench_p_cl1","bench_p_cl0","bench_p_f16_cl1","bench_p_f16_cl0"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d["roofline"]["classes"]
        print(f, d["ms_per_step"], "e2e", round(1e3*d["config"].get("B_per_gpu",64)*800/d["e2e"]["value"],3), {k:(round(v["ms_per_step"],3), v["launches_per_step"]) for k,v in c.items() if k in ("row_norm","dec.out_proj","dec.ffn_w2","enc.out_proj","enc.ffn_w2")})
    except Exception as e:
        print(f, "failed", e)
PY
