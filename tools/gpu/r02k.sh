#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_target_model.py -m gpu -q > $O/pytest_target.txt 2>&1
timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 3,2 > $O/gn_time_cfg2.txt 2>&1
timeout 300 python tools/gn_operator_time.py 5 80 80 30 54 10 3,2 > $O/gn_time_cfg3.txt 2>&1
timeout 300 python tools/gn_operator_time.py 10 32 32 45 80 10 3,2 > $O/gn_time_cfg5.txt 2>&1
timeout 300 python tools/gn_operator_time.py 1 20 80 30 54 5 3,2,1 > $O/gn_time_small.txt 2>&1
grep -E "passed|failed|^FAILED|Error|error" $O/pytest_target.txt | tail -8
head -n 2 $O/gn_time_cfg2.txt $O/gn_time_cfg3.txt $O/gn_time_cfg5.txt; cat $O/gn_time_small.txt
