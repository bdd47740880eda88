mkdir -p gpurun_out
FS2_PDL=1 timeout 600 python tools/pdl_check.py 2>&1 | tail -2
for pdl in 0 1 0 1; do
  FS2_PDL=$pdl python bench.py --gpus 1 --steps 20 --warmup 5 --modes "f16" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_h2_pdl$pdl.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_h2_pdl$pdl.json").read().strip().splitlines()[-1])
print("PDL=$pdl", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "busy", round(d["gpu_busy_ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d.get("modes",{}).items()})
PY
done
