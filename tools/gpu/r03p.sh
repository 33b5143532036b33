#!/bin/bash
O=gpurun_out/r03p; mkdir -p $O
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
grep -E "total kernel|memory_|build_stencil|pixel_weights|merge" $O/block_cfg2.md
