#!/bin/bash
O=gpurun_out/r04b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "stem" > $O/pytest_stem.log 2>&1; tail -12 $O/pytest_stem.log
