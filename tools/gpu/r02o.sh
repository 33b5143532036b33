#!/bin/bash
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_model.py -m gpu -q -s -x > $O/pytest_driver.txt 2>&1
FRTM_GRAPH_BLOCKS=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_driver.py -m gpu -q -x > $O/pytest_graph.txt 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_eager.json 2> $O/bench_eager.err
FRTM_GRAPH_BLOCKS=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_graph.json 2> $O/bench_graph.err
FRTM_GRAPH_BLOCKS=1 timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_graph3.json 2> $O/bench_graph3.err
tail -4 $O/pytest_driver.txt; tail -4 $O/pytest_graph.txt
for f in bench_eager bench_graph bench_graph3; do python -c "
import json,sys
d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], round(d['ms_per_step'],2))" 2>&1 | tail -1; tail -2 $O/$f.err; done
