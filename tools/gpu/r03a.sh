#!/bin/bash
O=gpurun_out/r03a; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_driver.py -x -q -m gpu > $O/pytest_driver.log 2>&1; tail -4 $O/pytest_driver.log
for c in 2 3; do timeout 600 python tools/prefetch_timeline.py $c > $O/timeline_$c.txt 2>&1; cat $O/timeline_$c.txt | tail -30; done
