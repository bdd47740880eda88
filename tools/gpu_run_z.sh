mkdir -p gpurun_out
FS2_LN_KH=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm" 2>&1 | tail -3 > gpurun_out/pytest_z_kh2.log; cat gpurun_out/pytest_z_kh2.log
for kh in 0 1 2 0 1; do
  FS2_LN_KH=$kh python bench.py --gpus 1 --steps 20 --warmup 5 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_z_kh$kh.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_z_kh$kh.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("KH=$kh", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items() if k in ("dec.out_proj","dec.ffn_w2","enc.out_proj","enc.ffn_w2","dec.ffn_w1_conv9","dec.qkv_proj")})
PY
done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm -s 15 -c 1 -f -o gpurun_out/qkv_x3_r2g python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_z1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm -s 24 -c 1 -f -o gpurun_out/postnet_x3_r2g python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_z2.log 2>&1
tail -2 gpurun_out/ncu_z1.log gpurun_out/ncu_z2.log
