#!/bin/bash
# end-of-round evidence (as r03s.sh), final code
O=gpurun_out/r04v; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
S=$(date +%s); python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench.py wall clock: $(( $(date +%s) - S )) s"
python tools/bench_brief.py $O/bench_default.json 2>&1 | head -14 | cut -c1-260; tail -2 $O/bench_default.err
M="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 ncu $M --log-file $O/block_cfg2.csv python tools/profile_step.py --what block --frames 33 > $O/block_cfg2.log 2>&1
timeout 900 ncu $M --log-file $O/block_cfg3.csv python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/block_cfg3.log 2>&1
python tools/summarize_launches.py $O/block_cfg2.csv --md "config 2 block" > $O/block_cfg2.md 2>&1
python tools/summarize_launches.py $O/block_cfg3.csv --md "config 3 block" > $O/block_cfg3.md 2>&1
head -3 $O/block_cfg2.md; grep -E "build_stencil|upsample_tapsum" $O/block_cfg2.md; head -3 $O/block_cfg3.md




