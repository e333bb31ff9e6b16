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
# full captures are summarised ON THE BOX (gpurun brings back at most 64 MiB): only the level-3 report itself is kept
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 sk=$3 ct=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $sk -c $ct -f -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1; tail -3 gpurun_out/$name.txt | cut -c1-200
}
cap r2_prof_tc_l3 'stats_t|attn_tc|fuse_tc|mask_tc' 37 7 $B
cap r2_prof_slot 'slot_|tav_|tscore' 70 10 $B; rm -f gpurun_out/r2_prof_slot.ncu-rep
cap r2_prof_misc 'fuse_count4|fuse_argmax4|mha_core|reduce_attn' 16 4 $B; rm -f gpurun_out/r2_prof_misc.ncu-rep
# the deformable-conv subnet next to the reference's own op on this GPU; ncu of the implicit-GEMM kernel (three layers, level 0)
timeout 300 python scripts/dcn_bench.py --ref --out gpurun_out/r2_dcn_bench.json > /dev/null 2> gpurun_out/dcn.err; tail -2 gpurun_out/dcn.err; python -c "import json; d=json.load(open('gpurun_out/r2_dcn_bench.json')); print('dcn', d['total_ms'], d.get('reference_total_ms'))"
cap r2_prof_dcn dcn_tc 0 3 python scripts/dcn_bench.py --levels 0; rm -f gpurun_out/r2_prof_dcn.ncu-rep
du -sh gpurun_out
