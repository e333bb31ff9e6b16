# round-2 GPU check: parity tests (with the printed error figures kept), smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rf ${PYTEST_ARGS:-} > gpurun_out/r2_pytest.log 2>&1
grep -E "rel|passed|failed|FAILED|Error|error|stage|fusion|mismatch|path=" gpurun_out/r2_pytest.log | grep -v "^E  " | tail -${TAILN:-70}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -4 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
