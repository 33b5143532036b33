#!/bin/bash
O=gpurun_out/r04k; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2.json 2> $O/bench2.err
python tools/bench_brief.py $O/bench2.json 2>&1 | head -2 | cut -c1-110; tail -2 $O/bench2.err
