timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --repeats 3 ${BENCH_ARGS:-} > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -3 gpurun_out/bq.err
python scripts/bench_summary.py gpurun_out/bq.json
