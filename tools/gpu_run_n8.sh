mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --modes "" --cpu-sample-batch 8 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_n8box_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --modes "" 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_n8box_n8.json
python - <<'PY'
import json
a=json.loads(open("gpurun_out/bench_n8box_n1.json").read().strip().splitlines()[-1])
b=json.loads(open("gpurun_out/bench_n8box_n8.json").read().strip().splitlines()[-1])
print("n1", a["value"], a["ms_per_step"], a["e2e"]["value"]); print("n8", b["value"], b["ms_per_step"], b["e2e"]["value"], "eff", b["value"]/(8*a["value"]), b["e2e"]["value"]/(8*a["e2e"]["value"]))
PY
