#!/bin/bash
O=gpurun_out/r04o; mkdir -p $O
T=$PWD/frtm_vos_b200/libfrtm_b200_timing.so
FRTM_B200_LIB=$T timeout 200 python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/timeline_cfg2.txt 2>&1
FRTM_B200_LIB=$T timeout 200 python tools/gn_operator_time.py 2 74 80 30 54 5 3 > $O/timeline_wave.txt 2>&1
for f in cfg2 wave; do echo "== $f"; grep "gm " $O/timeline_$f.txt | tail -9; grep -v "gm " $O/timeline_$f.txt | tail -1; done
