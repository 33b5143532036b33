#!/bin/bash
# ncu --set full of the CTA-pair conv kernel (N = 256) and the general kernel on the rn101 stage-3 3x3 and expand convs
O=gpurun_out/r02s; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 4 -o $O/conv_s3_3x3 python tools/conv_time.py --only "rn101 s3 3x3" --sels 1,3 --reps 2 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 4 -o $O/conv_s3_expand python tools/conv_time.py --only "rn101 s3 expand" --sels 1,3 --reps 2 > $O/ncu2.log 2>&1
ls -la $O; tail -3 $O/ncu1.log
