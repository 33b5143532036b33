#!/bin/bash
O=gpurun_out/r03v; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_target_model.py tests/test_gpu_driver.py -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
