mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_x.log; cat gpurun_out/pytest_x.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_x.log 2>&1; tail -4 gpurun_out/smoke_x.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_x_ref.json 2> gpurun_out/bench_x_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_x.err | grep '^{"metric"' > gpurun_out/bench_x.json
python bench.py --steps 10 --warmup 3 --workload c4 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_x_c4.json
python - <<'PY'
import json
for f in ["bench_x","bench_x_c4"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    c=d["roofline"]["classes"]
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "frac", d["roofline"]["frac"], d["clocks"]["reasons"], d["gpu_launches"], {k:round(v["ms_per_step"],3) for k,v in c.items()}, d.get("modes"))
print(open("gpurun_out/bench_x_ref.json").read()[:300])
PY
