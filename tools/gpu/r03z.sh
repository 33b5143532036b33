#!/bin/bash
O=gpurun_out/r03z; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
for v in 1 0 1 0; do
FRTM_BENCH_CHAIN=$v timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2_$v.json 2> $O/bench2.err
echo "chain=$v"; python tools/bench_brief.py $O/bench2_$v.json 2>&1 | head -2 | cut -c1-100; tail -1 $O/bench2.err
done
for v in 1 0; do
FRTM_BENCH_CHAIN=$v timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3_$v.json 2> $O/bench3.err
echo "chain=$v"; python tools/bench_brief.py $O/bench3_$v.json 2>&1 | head -2 | cut -c1-100; tail -1 $O/bench3.err
done
