#!/bin/bash
O=gpurun_out/r04e; mkdir -p $O
FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so timeout 300 python - > $O/stem_timing.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from frtm_vos_b200 import ops
g = torch.Generator().manual_seed(1)
img = torch.randint(0, 256, (8, 3, 480, 854), dtype=torch.uint8, generator=g).cuda()
w = torch.randn(64, 3, 7, 7, generator=g) / 12
pc = ops.pack_conv_tc(ops.stem_weight_as_1x1(w), None, device="cuda:0")
for _ in range(3):
    y = ops.stem_conv(img, pc)
torch.cuda.synchronize()
PY
grep "^stem" $O/stem_timing.log | tail -2
