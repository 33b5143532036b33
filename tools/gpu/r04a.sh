#!/bin/bash
O=gpurun_out/r04a; mkdir -p $O
for v in 0 0 0 0 1 1 1 1; do
FRTM_BENCH_CHAIN=$v timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/b.json 2> $O/b.err
python - <<PY
import json
d=json.loads(open("$O/b.json").read().strip().splitlines()[-1])
print("chain=$v value %.0f e2e %.0f no-overlap %.0f" % (d["value"], d["e2e"]["value"], d["pipeline"]["without_overlap"]["value"]))
PY
done
