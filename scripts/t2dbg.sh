timeout 150 python -m pytest tests -m gpu -q -s -x -k "slot_attention" 2>&1 | grep -E "illegal|rel err|passed|failed|Error" | head -20
timeout 250 python -m pytest tests -m gpu -q -s -k "whole or chain or stage" 2>&1 | grep -E "illegal|rel|passed|failed|Error|stage" | head -40
