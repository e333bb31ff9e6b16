# sweep side-stream x main-stream grid widths at a fixed number of clips in flight (value only)
for sc in ${SIDES:-48 64 80}; do for mc in ${MAINS:-48 64 80}; do
  SLOTVPS_SIDE_CTAS=$sc SLOTVPS_MAIN_CTAS=$mc timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --inflight ${INFLIGHT:-8} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side $sc main $mc', round(d['value'],1), round(d['ms_per_step'],4))"
done; done
