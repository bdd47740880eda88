mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for dbg in 0 1 2 3; do
  FS2_GEMM_DEBUG=$dbg python bench.py --gpus 1 --steps 10 --warmup 3 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_b2_dbg$dbg.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_b2_dbg$dbg.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("DBG=$dbg", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items() if k in ("enc.qkv_proj","dec.qkv_proj","predictor.tap_gemm","postnet.conv5","dec.ffn_w1_conv9","enc.ffn_w1_conv9","dec.embed_linear")})
PY
done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm -s 15 -c 1 -f -o gpurun_out/qkv_x3_r2h python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_b2.log 2>&1
tail -n 2 gpurun_out/ncu_b2.log
