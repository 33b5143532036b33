#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest_all.txt 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
grep -E "passed|failed|^FAILED|^P[145] " $O/pytest_all.txt | tail -12
tail -c 600 $O/bench.json; tail -3 $O/bench.err; tail -c 400 $O/bench_ref.json
