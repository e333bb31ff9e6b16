# BASELINE configs[2..4] through bench.py on one GPU
timeout 900 python bench.py --config viper --steps 12 --warmup 3 --repeats 3 --no-cpu-baseline > gpurun_out/r2_bench_viper_n1.json 2> gpurun_out/viper.err; tail -2 gpurun_out/viper.err; python scripts/bench_summary.py gpurun_out/r2_bench_viper_n1.json | head -3
timeout 900 python bench.py --config sweep --warmup 3 --repeats 2 --no-cpu-baseline > gpurun_out/r2_bench_sweep_n1.json 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; python scripts/bench_summary.py gpurun_out/r2_bench_sweep_n1.json | head -3
timeout 1200 python bench.py --config slots --steps 6 > gpurun_out/r2_bench_slots_n1.json 2> gpurun_out/slots.err; tail -2 gpurun_out/slots.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_slots_n1.json") if l.startswith("{")][-1])
for r in d["table"]: print(r["n_slots"], r["iterations"], round(r["ms_per_clip"],3), round(r["frames_per_s"],1), round(r["single_clip_ms"],3), r["launches_per_clip"])
PY
