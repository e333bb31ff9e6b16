# multi-GPU bench: N ranks of one node (run with gpurun --gpus N)
N=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --repeats 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
python scripts/bench_summary.py gpurun_out/bench_n$N.json | head -4
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json")); print(d["config"]["numa_bind"], d["config"]["sharding"])
PY
