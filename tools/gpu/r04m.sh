#!/bin/bash
# GPU call: phase timeline of the sliding-window operator kernel (timing build) at the config-2 / config-5 update shapes and for a single sample
O=gpurun_out/r04m; mkdir -p $O
T=$PWD/frtm_vos_b200/libfrtm_b200_timing.so
FRTM_B200_LIB=$T timeout 200 python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/timeline_cfg2.txt 2>&1
FRTM_B200_LIB=$T timeout 200 python tools/gn_operator_time.py 1 1 80 30 54 5 3 > $O/timeline_single.txt 2>&1
FRTM_B200_LIB=$T timeout 200 python tools/gn_operator_time.py 10 32 32 45 80 10 3 > $O/timeline_cfg5.txt 2>&1
for f in cfg2 single cfg5; do echo "== $f"; grep "gm timeline" $O/timeline_$f.txt | tail -2; grep -v "gm timeline" $O/timeline_$f.txt | tail -2; done
