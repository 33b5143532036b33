#!/bin/bash
O=gpurun_out/r03w; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_driver.py -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2.json 2> $O/bench2.err
python tools/bench_brief.py $O/bench2.json 2>&1 | head -2; tail -2 $O/bench2.err
timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3.json 2> $O/bench3.err
python tools/bench_brief.py $O/bench3.json 2>&1 | head -2; tail -2 $O/bench3.err
