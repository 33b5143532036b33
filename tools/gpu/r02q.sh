#!/bin/bash
# two ranks on one box: config 2 (weak scaling, one sequence per GPU) and config 4 (64 sequences sharded, strong scaling)
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 4 --steps 2 --warmup 3 > $O/bench_cfg4_n2.json 2> $O/bench_cfg4_n2.err
timeout 900 python bench.py --impl reference-cuda --steps 2 --warmup 1 > $O/bench_refcuda.json 2> $O/bench_refcuda.err
for f in bench_cfg2_n2 bench_cfg4_n2; do python tools/bench_brief.py $O/$f.json 2>&1 | head -3; tail -2 $O/$f.err; done
tail -c 500 $O/bench_refcuda.json
