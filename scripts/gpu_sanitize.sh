# compute-sanitizer memcheck + racecheck on small shapes (SURVEY.md section 4.1 "Sanitizers")
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests -m gpu -q -x \
     -k "(slot_attention and 100-16-32-True) or (single_stage and True-0) or fusion_b or test_mask_logits or tracker_golden_videos and track_c or track_head_scores and 22-37 or test_level_fuse and 64-128" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|hazard|Race" gpurun_out/sanitize_$tool.log | head -8
done
