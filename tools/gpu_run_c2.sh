mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for pf in 0 1 0 1; do
  FS2_GEMM_PREFETCH=$pf python bench.py --gpus 1 --steps 20 --warmup 5 --modes "f16" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_c2_pf$pf.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c2_pf$pf.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("PF=$pf", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items()}, {k:round(v["ms_per_step"],3) for k,v in d.get("modes",{}).items()})
PY
done
