#!/bin/bash
O=gpurun_out/r02h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_target_model.py -m gpu -q > $O/pytest_target.txt 2>&1
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 4 > $O/timeline_cfg2.txt 2>&1
timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 4,3,1 > $O/gn_time_cfg2.txt 2>&1
timeout 300 python tools/gn_operator_time.py 5 80 80 30 54 10 4,3,1 > $O/gn_time_cfg3.txt 2>&1
timeout 300 python tools/gn_operator_time.py 10 32 32 45 80 10 4,3,1 > $O/gn_time_cfg5.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_cl -s 8 -c 1 -o $O/gn_cl python tools/gn_operator_time.py 3 69 80 30 54 5 4 > $O/ncu.log 2>&1
grep -E "passed|failed|^FAILED|Error|error" $O/pytest_target.txt | tail -8
grep "cl timeline" $O/timeline_cfg2.txt | tail -2
head -n 1 $O/gn_time_cfg2.txt $O/gn_time_cfg3.txt $O/gn_time_cfg5.txt; tail -n 3 $O/gn_time_cfg2.txt
