#!/bin/bash
O=gpurun_out/r03d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_model.py -x -q -m gpu > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2.json 2> $O/bench2.err
python tools/bench_brief.py $O/bench2.json 2>&1 | head -3; tail -3 $O/bench2.err
