#!/bin/bash
O=gpurun_out/r02f; mkdir -p $O
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 4 > $O/timeline_cfg2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_target_model.py -m gpu -q > $O/pytest_target.txt 2>&1
grep "cl timeline" $O/timeline_cfg2.txt | tail -4
grep -E "passed|failed|^FAILED|Error|error" $O/pytest_target.txt | tail -8
