#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O
for c in 2 3; do timeout 600 python tools/step_breakdown.py $c > $O/breakdown_$c.txt 2>&1; tail -8 $O/breakdown_$c.txt; done
timeout 600 python tools/init_timeline.py > $O/init_timeline.txt 2>&1; tail -40 $O/init_timeline.txt
