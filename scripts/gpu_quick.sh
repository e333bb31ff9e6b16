# quick GPU loop: parity tests (stop at first failure) + bench summary
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_latest.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"])
kb=d["kernel_breakdown_ms_per_step"]
print({k:round(v["ms_per_step"],3) for k,v in sorted(kb.items(), key=lambda kv:-kv[1]["ms_per_step"])[:14] if k!="copy"})
PY
tail -2 gpurun_out/bench_latest.err
