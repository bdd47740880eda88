mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
for mg in 2 4; do FS2_LN_MG=$mg timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm" 2>&1 | tail -1 | sed "s/^/FS2_LN_MG=$mg /" | tee -a gpurun_out/pytest_final_flags.log; done
FS2_LN_KH=2 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_gemm_layernorm" 2>&1 | tail -1 | sed "s/^/FS2_LN_KH=2 /" | tee -a gpurun_out/pytest_final_flags.log
FS2_ATT_QT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention" 2>&1 | tail -1 | sed "s/^/FS2_ATT_QT=1 /" | tee -a gpurun_out/pytest_final_flags.log
FS2_PDL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or oracle" 2>&1 | tail -1 | sed "s/^/FS2_PDL=1 /" | tee -a gpurun_out/pytest_final_flags.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -4 gpurun_out/smoke_final.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_final.err | grep '^{"metric"' > gpurun_out/bench_final.json
python bench.py --steps 10 --warmup 3 --workload c4 --modes f16 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_final_c4.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_3xf16_final.csv python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_final_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tap_gemm -s 16 -c 1 -f -o gpurun_out/conv9_3x_final python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_final_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16_kernel -s 4 -c 1 -f -o gpurun_out/attn_x3_final python tools/profile_step.py --precision 3xf16 > gpurun_out/ncu_final_2.log 2>&1
python - <<'PY'
import json
for f in ["bench_final","bench_final_c4"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    c=d["roofline"]["classes"]
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"]["reasons"], d["gpu_launches"], {k:round(v["ms_per_step"],3) for k,v in c.items()}, {k:v["ms_per_step"] for k,v in d.get("modes",{}).items()})
print(open("gpurun_out/bench_final_ref.json").read()[:400])
PY
