#!/bin/bash
# GPU call: batched-row stencil build + 16-channel rank-1 finish — full GPU suite, launch list of a config-2 block
O=gpurun_out/r04r; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -5 $O/pytest.log
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
head -3 $O/block_cfg2.md; grep -E "upsample_tapsum|pyrup|build_stencil|merge_masks|rank1" $O/block_cfg2.md
