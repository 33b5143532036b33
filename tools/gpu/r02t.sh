#!/bin/bash
O=gpurun_out/r02t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "pair" > $O/pytest_pair.log 2>&1
tail -3 $O/pytest_pair.log
timeout 300 python tools/conv_time.py > $O/conv_time.md 2> $O/conv_time.err
cat $O/conv_time.md; tail -3 $O/conv_time.err
timeout 300 python tools/conv_time.py --no-flush > $O/conv_time_warm.md 2> $O/conv_time.err
cat $O/conv_time_warm.md; tail -3 $O/conv_time.err
