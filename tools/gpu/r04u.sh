#!/bin/bash
# GPU call: merged pool + gate — ops / model / driver tests, launch list of a config-2 block
O=gpurun_out/r04u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_driver.py -q -m gpu -x > $O/pytest.log 2>&1; tail -5 $O/pytest.log
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
head -3 $O/block_cfg2.md; grep -E "gap_|cab_" $O/block_cfg2.md
