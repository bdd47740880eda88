mkdir -p gpurun_out
for cfg in "2 1" "4 1" "1 1"; do
  set -- $cfg
  echo "=== FS2_LN_MG=$1 FS2_LN_AMC=$2"
  FS2_LN_MG=$1 FS2_LN_AMC=$2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm" 2>&1 | tail -3
  FS2_LN_MG=$1 FS2_LN_AMC=$2 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --modes "f16" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_t_mg$1_amc$2.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_t_mg$1_amc$2.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("bench mg$1 amc$2", d["ms_per_step"], "e2e", round(1e3*64*800/d["e2e"]["value"],3), {k:round(v["ms_per_step"],3) for k,v in c.items() if k in ("dec.out_proj","dec.ffn_w2","enc.out_proj","enc.ffn_w2","dec.ffn_w1_conv9")}, "f16", d["modes"]["f16"]["ms_per_step"])
PY
done
FS2_LN_MG=2 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
