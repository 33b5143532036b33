#!/bin/bash
O=gpurun_out/r03h; mkdir -p $O
timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3.json 2> $O/bench3.err
python tools/bench_brief.py $O/bench3.json 2>&1 | head -2; python -c "
import json;d=json.loads(open('$O/bench3.json').read().strip().splitlines()[-1]);print(d['roofline_conv']['ms_per_frame'])"; tail -3 $O/bench3.err
