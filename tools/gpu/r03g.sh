#!/bin/bash
O=gpurun_out/r03g; mkdir -p $O
timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3.json 2> $O/bench3.err
python tools/bench_brief.py $O/bench3.json 2>&1 | head -3; tail -3 $O/bench3.err
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg3.csv python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/block_cfg3.log 2>&1
python tools/summarize_launches.py $O/block_cfg3.csv --md "config 3 block" > $O/block_cfg3.md 2>&1
head -14 $O/block_cfg3.md
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
