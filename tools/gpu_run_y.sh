mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_y.log; cat gpurun_out/pytest_y.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_y.log 2>&1; tail -4 gpurun_out/smoke_y.log
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_y.err | grep '^{"metric"' > gpurun_out/bench_y.json
python bench.py --steps 10 --warmup 3 --workload c4 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_y_c4.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_3xf16_r2g.csv python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_y3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ln_cluster -s 8 -c 2 -f -o gpurun_out/gemm_ln_cl_x3_r2g python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_y1.log 2>&1
python - <<'PY'
import json
for f in ["bench_y","bench_y_c4"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    c=d["roofline"]["classes"]
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "frac", d["roofline"]["frac"], d["clocks"]["reasons"], d["gpu_launches"], {k:round(v["ms_per_step"],3) for k,v in c.items()}, d.get("modes"))
PY
