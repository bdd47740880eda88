mkdir -p gpurun_out
FS2_LN_XASYNC=1 timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm or (golden_teacher_forced and 3xtf32) or (golden_filelist_twin and 3xtf32)" 2>&1 | tail -1
for xa in 0 1; do
  FS2_LN_XASYNC=$xa timeout 60 python bench.py --gpus 1 --steps 10 --warmup 3 --modes "" --cpu-sample-batch 2 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_xa$xa.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_xa$xa.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("XA=$xa", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items() if k in ("enc.out_proj","enc.ffn_w2","dec.out_proj","dec.ffn_w2","dec.qkv_proj","dec.ffn_w1_conv9")})
PY
done
