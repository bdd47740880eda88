mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn" | tail -25 > gpurun_out/pytest_n.log
FS2_ATT_TRACE=1 timeout 300 python tools/attn_probe.py 2> gpurun_out/attn_trace_n.log | tail -3 > gpurun_out/attn_probe_n.log
timeout 300 python tools/attn_probe.py 32 2000 384 2 2>/dev/null | tail -3 >> gpurun_out/attn_probe_n.log
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_n.json
timeout 600 python bench.py --steps 20 --warmup 5 --precision f16 --modes '' 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_n_f16.json
tail -4 gpurun_out/pytest_n.log; cat gpurun_out/attn_probe_n.log; head -10 gpurun_out/attn_trace_n.log
python - <<'PY'
import json
for f in ["bench_n","bench_n_f16"]:
    d=json.load(open("gpurun_out/%s.json"%f)); c=d["roofline"]["classes"]
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in c.items()})
    if d.get("modes"): print({k:round(v["ms_per_step"],3) for k,v in d["modes"].items()})
PY
