#!/bin/bash
# full GPU test suite + bench lines (config 2 with configs 3 and 5 co-reported, no CPU arm) with the CTA-pair conv kernel
O=gpurun_out/r02v; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python tools/bench_brief.py $O/bench.json 2>&1 | head -20; tail -3 $O/bench.err
