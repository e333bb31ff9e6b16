#!/bin/bash
# slot-kernel iteration loop: single-stage teacher-forced parity, big golden + full-size parity, quick bench
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "test_single_stage_teacher_forced" 2>&1 | grep -E "slot_pre|slot_post|stage path|passed|failed|Error|error" | head -20
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "test_head_big_golden_teacher_forced or test_fullsize_head" 2>&1 | grep -E "passed|failed|Assertion|Error" | head
timeout 300 bash scripts/gpu_bench_quick.sh 2>&1 | head -40
