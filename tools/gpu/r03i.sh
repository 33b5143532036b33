#!/bin/bash
# eight ranks on one box: config 2 (weak scaling, one sequence per GPU) and config 4 (64 sequences x 4 objects sharded, strong scaling)
O=gpurun_out/r03i; mkdir -p $O
nproc > $O/cores.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 > $O/bench_cfg2_n8.json 2> $O/bench_cfg2_n8.err
python tools/bench_brief.py $O/bench_cfg2_n8.json 2>&1 | head -3; tail -2 $O/bench_cfg2_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 4 --steps 2 --warmup 3 > $O/bench_cfg4_n8.json 2> $O/bench_cfg4_n8.err
python tools/bench_brief.py $O/bench_cfg4_n8.json 2>&1 | head -3; tail -2 $O/bench_cfg4_n8.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_cfg2_n1.json 2> $O/bench_cfg2_n1.err
python tools/bench_brief.py $O/bench_cfg2_n1.json 2>&1 | head -2; cat $O/cores.txt
