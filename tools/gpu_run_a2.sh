mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_a2.log; cat gpurun_out/pytest_a2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_a2.log 2>&1; tail -4 gpurun_out/smoke_a2.log
for i in 1 2; do
python bench.py --gpus 1 --steps 20 --warmup 5 --modes "f16" 2> gpurun_out/bench_a2.err | grep '^{"metric"' > gpurun_out/bench_a2_$i.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_a2_$i.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("a2", round(d["ms_per_step"],3), d["value"], {k:round(v["ms_per_step"],3) for k,v in c.items()}, {k:v["ms_per_step"] for k,v in d.get("modes",{}).items()})
PY
done
