# ncu launch list (per-launch device time) and full captures of the tensor-core kernels (B200_PROFILING.md recipe)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 460 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stats_tc|attn_tc|fuse_tc' -s 20 -c 6 -o gpurun_out/prof_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
