#!/bin/bash
O=gpurun_out/r03u; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_target_model.py -x -q -m gpu > $O/pytest_tm.log 2>&1; tail -3 $O/pytest_tm.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2p -s 12 -c 8 -o $O/pairp_cfg3 python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/ncu_pairp.log 2>&1
ls -la $O | tail -3
