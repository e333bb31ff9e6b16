# round 2: ncu launch list + full captures (level-3 launches of the tensor-core kernels; slot-update and fusion kernels)
B="python bench.py --steps 1 --warmup 1 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/r2_launches.csv
# matching kernels per step, in order: f0 s0 a0 | f1c f1m s1 a1 s2 a2 | f2c f2m s3 a3 s4 a4 | f3c f3m s5 a5 s6 a6 | m  (22)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stats_t|attn_tc|fuse_tc|mask_tc' -s 37 -c 7 -o gpurun_out/r2_prof_tc_l3 $B > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slot_p|fuse_count4|fuse_argmax4' -s 19 -c 6 -o gpurun_out/r2_prof_slot_fusion $B > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
