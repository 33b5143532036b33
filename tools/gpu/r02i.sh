#!/bin/bash
O=gpurun_out/r02i; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest_all.txt 2>&1
for c in 2 3 5; do timeout 600 python tools/step_breakdown.py $c > $O/breakdown_cfg$c.txt 2>&1; done
timeout 300 python tools/gn_operator_time.py 3 69 80 30 54 5 4,3,2 > $O/gn_time_cfg2.txt 2>&1
timeout 300 python tools/init_timeline.py > $O/init_timeline.txt 2>&1
grep -E "passed|failed|^FAILED|^P[145] " $O/pytest_all.txt | tail -12
tail -n 1 $O/breakdown_cfg2.txt $O/breakdown_cfg3.txt $O/breakdown_cfg5.txt
head -n 3 $O/gn_time_cfg2.txt; tail -n 22 $O/init_timeline.txt
