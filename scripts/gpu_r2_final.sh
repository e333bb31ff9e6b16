#!/bin/bash
# round-2 evidence run on one B200: parity log, smoke, bench lines (own arm, reference arm, configs 2-4), ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rf > gpurun_out/r2_pytest.log 2>&1; tail -2 gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err; python scripts/bench_summary.py gpurun_out/r2_bench_n1.json | head -4
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; tail -c 600 gpurun_out/r2_bench_reference_arm.json
timeout 600 python bench.py --inflight 1 --no-cpu-baseline > gpurun_out/r2_bench_n1_inflight1.json 2>/dev/null; python scripts/bench_summary.py gpurun_out/r2_bench_n1_inflight1.json | head -3
bash scripts/gpu_configs.sh
bash scripts/gpu_launchlist.sh
B="python bench.py --steps 1 --warmup 1 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-graph"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stats_t|attn_tc|fuse_tc|mask_tc' -s 37 -c 7 -f -o gpurun_out/r2_prof_tc_l3 $B > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slot_|tav_|tscore' -s 70 -c 10 -f -o gpurun_out/r2_prof_slot $B > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fuse_count4|fuse_argmax4|mha_core|reduce_attn' -s 16 -c 4 -f -o gpurun_out/r2_prof_misc $B > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
