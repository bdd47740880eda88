mkdir -p gpurun_out
for dbg in 0 4 0 4; do
  FS2_GEMM_DEBUG=$dbg python bench.py --gpus 1 --steps 10 --warmup 3 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_f2_dbg$dbg.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_f2_dbg$dbg.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("DBG=$dbg", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items() if k in ("enc.qkv_proj","dec.qkv_proj","dec.attention","dec.ffn_w1_conv9")})
PY
done
