mkdir -p gpurun_out
for C in peer_store none peer_copy; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --collective $C 2> gpurun_out/bench_j_n2_$C.err | grep '^{"metric"' > gpurun_out/bench_j_n2_$C.json
done
python - <<'PY'
import json
for f in ["peer_store","none","peer_copy"]:
    try:
        d=json.load(open("gpurun_out/bench_j_n2_%s.json"%f)); print(f, round(d["value"]/1e6,3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]/1e6,3), round(d["e2e"]["ms_per_step"],3), d.get("collective_note"))
    except Exception as e: print(f, "ERR", e)
PY
tail -c 1500 gpurun_out/bench_j_n2_peer_store.err
