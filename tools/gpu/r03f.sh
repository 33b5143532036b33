#!/bin/bash
O=gpurun_out/r03f; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu > $O/pytest_conv.log 2>&1; tail -6 $O/pytest_conv.log
timeout 400 python tools/conv_time.py --only rn101 > $O/conv_time.md 2> $O/conv_time.err
cat $O/conv_time.md; tail -3 $O/conv_time.err
