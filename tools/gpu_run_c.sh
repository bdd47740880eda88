mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -60 > gpurun_out/pytest_c.log
for cfg in "1 0" "1 1" "1 2" "1 4" "1 7" "0 0"; do set -- $cfg; FS2_ATT_X2=$1 FS2_ATT_DEBUG=$2 python tools/attn_probe.py 2>&1 | grep -E "X2=" ; done > gpurun_out/attn_probe_c.log 2>&1
FS2_ATT_X2=1 python tools/attn_probe.py 32 2000 384 2 2>&1 | grep -E "X2=" >> gpurun_out/attn_probe_c.log
tail -25 gpurun_out/pytest_c.log; cat gpurun_out/attn_probe_c.log
