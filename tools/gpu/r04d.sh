#!/bin/bash
O=gpurun_out/r04d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "stem" > $O/pytest_stem.log 2>&1; tail -3 $O/pytest_stem.log
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
grep -E "total kernel|stem|maxpool" $O/block_cfg2.md
