import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step", "clocks", "model_tflops", "dtype")})
print("e2e", d.get("e2e")); print("cpu", d.get("cpu_baseline"))
r = d.get("roofline")
if r:
    print(r["kernel"], "achieved", round(r["achieved"], 1), "frac", round(r["frac"], 3), "share", round(r["share_of_step"], 3))
    for k, v in r["classes"].items():
        print(f"{k:24s} {v['ms_per_step']:8.3f} ms  {v['launches_per_step']:3d}  {(v['tflops'] or 0):8.1f} TF  {(v['gbs'] or 0):8.0f} GB/s")
