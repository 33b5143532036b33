#!/bin/bash
O=gpurun_out/r04l; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
grep -E "total kernel|memory_" $O/block_cfg2.md
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2.json 2> $O/bench2.err
python tools/bench_brief.py $O/bench2.json 2>&1 | head -2 | cut -c1-110; tail -2 $O/bench2.err
