# sweep side-stream CTA count x clips in flight (value only)
for f in ${INFLIGHTS:-4 6}; do for c in ${CTAS:-64 80 96}; do
  SLOTVPS_SIDE_CTAS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --inflight $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight $f ctas $c', round(d['value'],1), round(d['ms_per_step'],4))"
done; done
