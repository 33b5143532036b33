#!/bin/bash
# ncu --set full of the elementwise kernels of a warm config-2 track block
O=gpurun_out/r03k; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"upsample_tapsum|pyrup|build_stencil|rank1_finish|stem_patches|merge_masks|resize_bilinear|gap_stage1|maxpool|cab_apply|split_kernel|pixel_weights|corr3x3" -c 40 -o $O/glue_cfg2 python tools/profile_step.py --what block --frames 33 > $O/ncu_glue.log 2>&1
ls -la $O; tail -2 $O/ncu_glue.log
