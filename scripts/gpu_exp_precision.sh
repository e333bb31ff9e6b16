timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_slot_attention_teacher_forced or test_head_big_golden_teacher_forced or test_stale_workspace or test_single_stage" 2>&1 | grep -E "slot_attention|teacher-forced|stage path|passed|failed|Assertion"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
