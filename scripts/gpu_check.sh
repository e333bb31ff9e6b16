# full GPU check: parity tests, smoke, bench (run on the B200 box via gpurun)
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel|passed|failed|Error|stage|fusion_" | tail -50
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 500 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -c 3500 gpurun_out/bench_latest.json; tail -3 gpurun_out/bench_latest.err
