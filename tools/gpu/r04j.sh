#!/bin/bash
O=gpurun_out/r04j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_model.py -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
grep -E "total kernel|conv_tc3_kernel<32" $O/block_cfg2.md
