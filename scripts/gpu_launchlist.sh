#!/bin/bash
# ncu launch list of one bench step (single clip, eager launches): true per-kernel durations, cold-cache and serialised
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/r2_launches.csv
