for dbg in 7 15 23 31; do FS2_LN_DEBUG=$dbg timeout 200 python tools/ln_time.py 2>&1 | grep "us$"; done
