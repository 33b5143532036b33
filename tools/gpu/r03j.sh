#!/bin/bash
O=gpurun_out/r03j; mkdir -p $O
S=$(date +%s)
python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench.py wall clock: $(( $(date +%s) - S )) s"
python tools/bench_brief.py $O/bench_default.json 2>&1 | head -40; tail -3 $O/bench_default.err
