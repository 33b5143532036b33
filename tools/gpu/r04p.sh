#!/bin/bash
# GPU call: ncu source-level capture of the sliding-window operator kernel at the config-2 update shape
O=gpurun_out/r04p; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_apply_mma -s 8 -c 1 -o $O/gn_mma python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/ncu.log 2>&1
tail -2 $O/ncu.log; ls -la $O
