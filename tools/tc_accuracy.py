"""Accuracy probe of the tensor-core conv: signed relative error vs fp64 for zero-mean and for all-positive operands
(exposes the rounding mode of the tensor-core accumulator)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frtm_vos_b200 import ops
DEV = "cuda:0"
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
nchw = lambda x: x.permute(0, 3, 1, 2).contiguous()
for cin in (64, 256, 1024):
    for positive in (False, True):
        g = torch.Generator().manual_seed(cin)
        x = torch.randn(1, cin, 30, 54, generator=g)
        w = torch.randn(64, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        if positive:
            x, w = x.abs(), w.abs()
        ref = F.conv2d(x.double(), w.double(), None, 1, 1)
        ref32 = F.conv2d(x, w, None, 1, 1).double()
        y_tc = nchw(ops.conv2d_tc(ops.split_f16(nhwc(x).to(DEV)), ops.pack_conv_tc(w, None, device=DEV))["y"].cpu()).double()
        y_si = nchw(ops.conv2d(nhwc(x).to(DEV), ops.pack_conv(w, None, device=DEV)).cpu()).double()
        den = ref.abs().clamp_min(1e-3)
        for name, y in (("tc", y_tc), ("simt", y_si), ("cpu32", ref32)):
            rel = (y - ref) / den
            print("K=%5d %s %-5s mean signed rel %+.2e  rms %.2e  max %.2e" % (cin * 9, "pos " if positive else "zero", name,
                  rel.mean().item(), rel.pow(2).mean().sqrt().item(), rel.abs().max().item()))
