SLOTVPS_SLOT_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_single_stage_teacher_forced" 2>&1 | grep -E "slot_pre|slot_post|stage path|passed|failed|Error" | head -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_head_big_golden_teacher_forced or test_fullsize_head" 2>&1 | grep -E "passed|failed|Assertion|Error" | head
bash scripts/gpu_bench_quick.sh 2>&1 | head -14
