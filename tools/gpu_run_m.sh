mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention or golden_teacher or oracle_long or edge or random_shapes" 2>&1 | grep -v "Warning\|warnings.warn" | tail -20 > gpurun_out/pytest_m.log
FS2_ATT_TRACE=1 timeout 300 python tools/attn_probe.py 2> gpurun_out/attn_trace_m.log | tail -3 > gpurun_out/attn_probe_m.log
timeout 300 python tools/attn_probe.py 32 2000 384 2 2>/dev/null | tail -3 >> gpurun_out/attn_probe_m.log
timeout 600 python bench.py --steps 20 --warmup 5 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_m.json
tail -4 gpurun_out/pytest_m.log; cat gpurun_out/attn_probe_m.log; head -10 gpurun_out/attn_trace_m.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_m.json")); c=d["roofline"]["classes"]
print(round(d["ms_per_step"],3), "attn", round(c["dec.attention"]["ms_per_step"],3), "enc attn", round(c["enc.attention"]["ms_per_step"],3), "f16:", round(d["modes"]["f16"]["ms_per_step"],3))
PY
