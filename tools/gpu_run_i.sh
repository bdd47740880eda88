mkdir -p gpurun_out
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_i_n${N}.err | grep '^{"metric"' > gpurun_out/bench_i_n${N}.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --collective all_gather 2> gpurun_out/bench_i_n8_ag.err | grep '^{"metric"' > gpurun_out/bench_i_n8_ag.json
python bench.py --gpus 1 --steps 20 --warmup 5 --modes '' 2> gpurun_out/bench_i_n1.err | grep '^{"metric"' > gpurun_out/bench_i_n1.json
python - <<'PY'
import json
for f in ["n1","n4","n8","n8_ag"]:
    try:
        d=json.load(open("gpurun_out/bench_i_%s.json"%f)); print(f, round(d["value"]/1e6,3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]/1e6,3), d.get("collective_note"))
    except Exception as e: print(f, "ERR", e)
PY
tail -c 600 gpurun_out/bench_i_n8.err
