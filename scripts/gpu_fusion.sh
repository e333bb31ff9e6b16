timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fusion or tracker or postprocess or whole_clip" 2>&1 | grep -E "fusion|passed|failed|Assertion|Error" | cut -c1-300 | head -30
bash scripts/gpu_bench_quick.sh 2>&1 | grep -E "value|single|panoptic_fusion|fuse_"
