#!/bin/bash
# GPU call: phase timeline + ncu capture of the single-pass operator kernel, then the GPU test suite
O=gpurun_out/r02b; mkdir -p $O
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/timeline_cfg2.txt 2>&1
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/gn_operator_time.py 1 1 80 30 54 5 3 > $O/timeline_single.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_mma -s 8 -c 2 -o $O/gn_mma python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/ncu.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest_all.txt 2>&1
grep "gm timeline" $O/timeline_cfg2.txt | tail -3; grep -v "gm timeline" $O/timeline_cfg2.txt | tail -3
grep "gm timeline" $O/timeline_single.txt | tail -2; grep -v "gm timeline" $O/timeline_single.txt | tail -2
tail -3 $O/ncu.log; grep -E "passed|failed|^FAILED|^P[145] " $O/pytest_all.txt | tail -20
