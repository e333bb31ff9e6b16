# multi-GPU runs: N ranks of one node (run with gpurun --gpus N); CONFIGS = list of bench configs
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR scripts/h2d_probe_ranks.py 2>/dev/null | grep ranks > gpurun_out/r2_h2d_ceiling_n$N.json; cat gpurun_out/r2_h2d_ceiling_n$N.json
for cfg in ${CONFIGS:-clip}; do
  extra=""; [ "$cfg" = "sweep" ] && extra="--repeats 2"; [ "$cfg" = "viper" ] && extra="--steps 12 --repeats 3"; [ "$cfg" = "clip" ] && extra="--steps 20 --repeats 5"
  $TR bench.py --gpus $N --config $cfg --warmup 3 $extra > gpurun_out/r2_bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err
  tail -2 gpurun_out/bench_${cfg}_n$N.err | cut -c1-300
  python scripts/bench_summary.py gpurun_out/r2_bench_${cfg}_n$N.json 2>/dev/null | head -2
done
