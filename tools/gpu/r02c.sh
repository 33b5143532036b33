#!/bin/bash
# GPU call: restructured single-pass operator — timeline, timing at the three config shapes, target-model tests, breakdowns
O=gpurun_out/r02c; mkdir -p $O
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/timeline_cfg2.txt 2>&1
timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 3,2 > $O/gn_time_cfg2.txt 2>&1
timeout 300 python tools/gn_operator_time.py 5 80 80 30 54 10 3,2 > $O/gn_time_cfg3.txt 2>&1
timeout 300 python tools/gn_operator_time.py 10 32 32 45 80 10 3,2 > $O/gn_time_cfg5.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_target_model.py tests/test_gpu_driver.py -m gpu -q > $O/pytest_target.txt 2>&1
for c in 2 3 5; do timeout 600 python tools/step_breakdown.py $c > $O/breakdown_cfg$c.txt 2>&1; done
grep "gm timeline" $O/timeline_cfg2.txt | tail -2
tail -2 $O/gn_time_cfg2.txt $O/gn_time_cfg3.txt $O/gn_time_cfg5.txt
grep -E "passed|failed|^FAILED" $O/pytest_target.txt | tail; tail -2 $O/breakdown_cfg*.txt
