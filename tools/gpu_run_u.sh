FS2_LN_MG=2 timeout 120 python tools/ln_probe.py 1024 384 384 f16 2>&1 | grep -c "nan 0"
