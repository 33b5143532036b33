#!/bin/bash
# ncu --set full captures of the default operator kernel at the three BASELINE update shapes (DRAM traffic, pipe utilisation)
O=gpurun_out/r02n; mkdir -p $O
N="--set full --clock-control none --import-source on -k regex:gn_apply_mma -s 8 -c 2"
timeout 600 ncu $N -o $O/gn_cfg2 python tools/gn_operator_time.py 3 69 80 30 54 5 3 > $O/ncu_cfg2.log 2>&1
timeout 600 ncu $N -o $O/gn_cfg3 python tools/gn_operator_time.py 5 80 80 30 54 10 3 > $O/ncu_cfg3.log 2>&1
timeout 600 ncu $N -o $O/gn_cfg5 python tools/gn_operator_time.py 10 32 32 45 80 10 3 > $O/ncu_cfg5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gn_apply_cl -s 8 -c 1 -o $O/gn_cl_cfg2 python tools/gn_operator_time.py 3 69 80 30 54 5 4 > $O/ncu_cl.log 2>&1
# conv kernels of a config-3 block: one full capture of the general kernel's main rn101 shapes
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc_kernel -c 12 -o $O/conv_cfg3 python tools/profile_step.py --arch resnet101 --objects 5 --full --what block --frames 33 > $O/ncu_conv.log 2>&1
ls -la $O
