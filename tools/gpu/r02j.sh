#!/bin/bash
# launch lists of one warm track block (configs 2 and 3) and of one object initialisation; compute-sanitizer runs
O=gpurun_out/r02j; mkdir -p $O
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
timeout 900 ncu $M --log-file $O/block_cfg3.csv python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/block_cfg3.log 2>&1
timeout 900 ncu $M --log-file $O/init_cfg3.csv python tools/profile_step.py --arch resnet101 --objects 1 --full --what init --frames 9 > $O/init_cfg3.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
python tools/summarize_launches.py $O/block_cfg3.csv --md "config 3 block" > $O/block_cfg3.md 2>&1
python tools/summarize_launches.py $O/init_cfg3.csv --md "config 3 init" > $O/init_cfg3.md 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gn_operator_time.py 2 5 6 30 54 2 4,3,2 > $O/memcheck_gn.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gn_operator_time.py 2 5 6 30 54 2 4,3 > $O/racecheck_gn.txt 2>&1
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_step.py --what block --frames 17 > $O/memcheck_block.txt 2>&1
head -30 $O/block_cfg3.md; head -16 $O/init_cfg3.md
tail -3 $O/memcheck_smoke.txt $O/memcheck_gn.txt $O/racecheck_gn.txt $O/memcheck_block.txt
