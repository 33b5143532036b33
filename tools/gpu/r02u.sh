#!/bin/bash
O=gpurun_out/r02u; mkdir -p $O
for shp in "rn101 s3 expand" "rn101 s2 expand" "rn101 s2 3x3"; do
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python tools/conv_time.py --only "$shp" --sels 4 --reps 1 > $O/timing.log 2>&1
echo "== $shp"; grep "^tc2p" $O/timing.log | tail -4
done
