mkdir -p gpurun_out
for C in none peer_store peer_copy; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 --collective $C 2> gpurun_out/bench_k_n8_$C.err | grep '^{"metric"' > gpurun_out/bench_k_n8_$C.json
done
python bench.py --gpus 1 --steps 20 --warmup 5 --modes '' 2> /dev/null | grep '^{"metric"' > gpurun_out/bench_k_n1.json
python - <<'PY'
import json
for f in ["n1","n8_none","n8_peer_store","n8_peer_copy"]:
    try:
        d=json.load(open("gpurun_out/bench_k_%s.json"%f)); print(f, round(d["value"]/1e6,3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]/1e6,3), round(d["e2e"]["ms_per_step"],3), d.get("collective_note"))
    except Exception as e: print(f, "ERR", e)
PY
