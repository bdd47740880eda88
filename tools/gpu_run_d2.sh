mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_robustness.py -x -q -m gpu -k "sharded or two_devices" 2>&1 | tail -15 > gpurun_out/pytest_d2_2gpu.log
cat gpurun_out/pytest_d2_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_d2_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_d2_n2.json").read().strip().splitlines()[-1])
print("n2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("collective"))
PY
