#!/bin/bash
O=gpurun_out/r03o; mkdir -p $O
for v in 1 0 1 0; do
FRTM_INSERT_BLOCK=$v timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2_$v.json 2> $O/bench2.err
echo "insert_block=$v"; python tools/bench_brief.py $O/bench2_$v.json 2>&1 | head -1 | cut -c1-90
done
for v in 1 0; do
FRTM_INSERT_BLOCK=$v timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3_$v.json 2> $O/bench3.err
echo "insert_block=$v"; python tools/bench_brief.py $O/bench3_$v.json 2>&1 | head -1 | cut -c1-90
done
