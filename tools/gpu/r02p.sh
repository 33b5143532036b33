#!/bin/bash
O=gpurun_out/r02p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest_all.txt 2>&1
timeout 1500 python bench.py --steps 8 --warmup 3 > $O/bench.json 2> $O/bench.err
grep -E "passed|failed|^FAILED|^P[145] |ytvos" $O/pytest_all.txt | tail -12
tail -c 300 $O/bench.json; tail -3 $O/bench.err
