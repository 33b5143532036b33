#!/bin/bash
O=gpurun_out/r02w; mkdir -p $O
timeout 300 python tools/conv_time.py --only rn101 --sels 1,0 > $O/conv_time.md 2> $O/conv_time.err
cat $O/conv_time.md; tail -3 $O/conv_time.err
timeout 300 python tools/tc_accuracy.py 2>&1 | grep " tc " 
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench3.json 2> $O/bench3.err
python tools/bench_brief.py $O/bench3.json 2>&1 | head -20; tail -3 $O/bench3.err
