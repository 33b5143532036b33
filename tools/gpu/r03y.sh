#!/bin/bash
O=gpurun_out/r03y; mkdir -p $O
for v in 1 0 1 0; do
FRTM_BENCH_STREAM_D2H=$v timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2_$v.json 2> $O/bench2.err
echo "stream_d2h=$v"; python tools/bench_brief.py $O/bench2_$v.json 2>&1 | head -2 | cut -c1-100
done
