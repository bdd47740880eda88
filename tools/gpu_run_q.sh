mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention" 2>&1 | tail -8 > gpurun_out/pytest_q_att.log
cat gpurun_out/pytest_q_att.log
timeout 300 python tools/attn_probe.py 2>&1 | grep -E "kernel only" > gpurun_out/attn_probe_q.log
cat gpurun_out/attn_probe_q.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_q.log
cat gpurun_out/pytest_q.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --modes "f16" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_q.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
c=d["roofline"]["classes"]
print("bench_q", d["ms_per_step"], "e2e", round(1e3*64*800/d["e2e"]["value"],3), {k:round(v["ms_per_step"],3) for k,v in c.items()}, d.get("modes"))
PY
