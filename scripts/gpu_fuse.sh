timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_level_fuse or test_head_chain or test_fullsize_head or test_folded or test_explicit_pos or test_stale or test_whole_clip or test_config4 or test_viper" 2>&1 | tail -3
bash scripts/gpu_bench_quick.sh 2>&1 | grep -E "value|single|level_fusion|fuse_tc|mask_logits"
