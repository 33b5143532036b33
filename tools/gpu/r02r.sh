#!/bin/bash
# CTA-pair conv kernel: parity at the production shapes, then timings against the general kernel
O=gpurun_out/r02r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "pair or matches_fp32 or stride2" > $O/pytest_pair.log 2>&1
tail -15 $O/pytest_pair.log
timeout 300 python tools/conv_time.py > $O/conv_time.md 2> $O/conv_time.err
cat $O/conv_time.md; tail -3 $O/conv_time.err
