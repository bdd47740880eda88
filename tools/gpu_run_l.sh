mkdir -p gpurun_out
FS2_ATT_QT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention" 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/pytest_l_qt.log
FS2_ATT_QT=1 FS2_ATT_TRACE=1 timeout 300 python tools/attn_probe.py 2> gpurun_out/attn_trace_l_qt.log | tail -3 > gpurun_out/attn_probe_l.log
FS2_ATT_QT=1 timeout 300 python tools/attn_probe.py 32 2000 384 2 2>/dev/null | tail -3 >> gpurun_out/attn_probe_l.log
echo "--- QT=0" >> gpurun_out/attn_probe_l.log
FS2_ATT_QT=0 timeout 300 python tools/attn_probe.py 2>/dev/null | tail -3 >> gpurun_out/attn_probe_l.log
FS2_ATT_QT=0 timeout 300 python tools/attn_probe.py 32 2000 384 2 2>/dev/null | tail -3 >> gpurun_out/attn_probe_l.log
FS2_ATT_QT=1 timeout 600 python bench.py --steps 20 --warmup 5 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_l_qt1.json
FS2_ATT_QT=0 timeout 600 python bench.py --steps 20 --warmup 5 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_l_qt0.json
tail -6 gpurun_out/pytest_l_qt.log; cat gpurun_out/attn_probe_l.log
python - <<'PY'
import json
for f in ["qt1","qt0"]:
    try:
        d=json.load(open("gpurun_out/bench_l_%s.json"%f)); c=d["roofline"]["classes"]
        print(f, round(d["ms_per_step"],3), "attn", round(c["dec.attention"]["ms_per_step"],3), "enc attn", round(c["enc.attention"]["ms_per_step"],3), "f16:", round(d["modes"]["f16"]["ms_per_step"],3))
    except Exception as e: print(f, "ERR", e)
PY
