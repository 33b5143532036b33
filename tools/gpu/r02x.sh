#!/bin/bash
# launch lists of one warm track block (configs 2 and 3) with the CTA-pair conv kernel
O=gpurun_out/r02x; mkdir -p $O
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg3.csv python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/block_cfg3.log 2>&1
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
python tools/summarize_launches.py $O/block_cfg3.csv --md "config 3 block" > $O/block_cfg3.md 2>&1
head -24 $O/block_cfg3.md
