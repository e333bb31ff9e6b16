for m in 1 2 3 4; do
  timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --inflight $m > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_m$m.json'))
    print("inflight=$m value %.1f ms/step %.3f | e2e %.1f (%.3f ms) | single %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['config'].get('single_clip_in_flight_ms_per_step')))
except Exception as e:
    print("inflight=$m failed", e); print(open('gpurun_out/bench_m$m.err').read()[-1500:])
PY
done
