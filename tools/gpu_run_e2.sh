mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
  python bench.py --gpus 1 --steps 20 --warmup 5 --modes "f16" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_e2_$i.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e2_$i.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("e2", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items()}, {k:round(v["ms_per_step"],3) for k,v in d.get("modes",{}).items()})
PY
done
