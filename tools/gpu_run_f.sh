mkdir -p gpurun_out
python -m pytest tests/test_gpu_robustness.py -m gpu -q --timeout 900 -p no:cacheprovider -k "two_devices or sharded_two" 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/pytest_f.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_f_n2_peer.json 2> gpurun_out/bench_f_n2_peer.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --collective all_gather > gpurun_out/bench_f_n2_ag.json 2> gpurun_out/bench_f_n2_ag.err
tail -15 gpurun_out/pytest_f.log; tail -c 800 gpurun_out/bench_f_n2_peer.err; head -c 600 gpurun_out/bench_f_n2_peer.json
