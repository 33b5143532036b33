#!/bin/bash
# end of round: eight ranks on one box, config 2 (weak scaling) against one GPU of the same box
O=gpurun_out/r04f; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 > $O/bench_cfg2_n8.json 2> $O/bench_cfg2_n8.err
python tools/bench_brief.py $O/bench_cfg2_n8.json 2>&1 | head -2 | cut -c1-200; tail -2 $O/bench_cfg2_n8.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_cfg2_n1.json 2> $O/bench_cfg2_n1.err
python tools/bench_brief.py $O/bench_cfg2_n1.json 2>&1 | head -2 | cut -c1-200
