#!/bin/bash
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ytvos.py -m gpu -q -s > $O/pytest_ytvos.txt 2>&1
tail -25 $O/pytest_ytvos.txt
