mkdir -p gpurun_out
PROBE_MODES=3xf16 FS2_ATT_TRACE=1 timeout 300 python tools/attn_probe.py > gpurun_out/attn_trace_x3.log 2>&1
cat gpurun_out/attn_trace_x3.log | tail -24
