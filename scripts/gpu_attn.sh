timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_slot_attention_teacher_forced or test_config5_slot_sweep or test_stale_workspace or test_slot_groups" 2>&1 | grep -E "path=0|passed|failed|Assertion|Error|N=" | cut -c1-200 | head -30
bash scripts/gpu_bench_quick.sh 2>&1 | grep -E "value|single|attention_contraction|attn_tc|stats_tc"
