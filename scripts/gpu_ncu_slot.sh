# ncu full capture of the slot-update kernels (one launch each) on a small single-stage case
export SLOTVPS_NO_OVERLAP=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slot_p -c 3 -o gpurun_out/prof_slot python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_single_stage_teacher_forced" > gpurun_out/ncu_slot.log 2>&1
tail -3 gpurun_out/ncu_slot.log
