#!/bin/bash
# evidence for profiles/: ncu --set full of the CTA-pair conv kernel inside a warm config-3 block, compute-sanitizer over the pair kernel tests
O=gpurun_out/r03e; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -s 30 -c 9 -o $O/pair_cfg3 python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/ncu_pair.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "pair" > $O/memcheck_pair.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "pair and 256-256-3-1-hw1-1" > $O/racecheck_pair.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_driver.py -x -q -m gpu -k "prefetched" > $O/memcheck_prefetch.txt 2>&1
ls -la $O; tail -4 $O/memcheck_pair.txt $O/racecheck_pair.txt $O/memcheck_prefetch.txt
