#!/bin/bash
# GPU call: default bench line with the NVML clock sampler
O=gpurun_out/r04t; mkdir -p $O
S=$(date +%s); python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench.py wall clock: $(( $(date +%s) - S )) s"
python tools/bench_brief.py $O/bench_default.json 2>&1 | head -14 | cut -c1-260; tail -2 $O/bench_default.err
python -c "
import json; d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print('clocks', d['clocks']); print([ (k, v.get('clocks')) for k, v in (d.get('also') or {}).items()] if isinstance(d.get('also'), dict) else '')
"
