#!/bin/bash
# two ranks on one box: config 2 (weak scaling, one sequence per GPU), then config 4 (64 sequences sharded)
O=gpurun_out/r03c; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err
python tools/bench_brief.py $O/bench_cfg2_n2.json 2>&1 | head -3; tail -2 $O/bench_cfg2_n2.err
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_cfg2_n1.json 2> $O/bench_cfg2_n1.err
python tools/bench_brief.py $O/bench_cfg2_n1.json 2>&1 | head -3; tail -2 $O/bench_cfg2_n1.err
