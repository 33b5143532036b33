"""Times frtm_conv2d_tc kernels on the wide-conv shapes of the ResNet-101 / ResNet-18 backbones (CUDA events, L2 flushed
between launches): general tile kernel (kernel_select 1) against the CTA-pair kernel (2: N = 128, 3: N = 256, 0: library's
choice).  python tools/conv_time.py [--frames 8]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frtm_vos_b200 import ops  # noqa: E402

SHAPES = [  # cin, cout, k, stride, (h, w) of the INPUT, residual
    ("rn101 s2 reduce", 512, 128, 1, 1, (60, 107), False),
    ("rn101 s2 3x3", 128, 128, 3, 1, (60, 107), False),
    ("rn101 s2 expand", 128, 512, 1, 1, (60, 107), True),
    ("rn101 s3 reduce", 1024, 256, 1, 1, (30, 54), False),
    ("rn101 s3 3x3", 256, 256, 3, 1, (30, 54), False),
    ("rn101 s3 expand", 256, 1024, 1, 1, (30, 54), True),
    ("rn101 s3 down", 512, 1024, 1, 2, (60, 107), False),
    ("rn101 s4 reduce", 2048, 512, 1, 1, (15, 27), False),
    ("rn101 s4 3x3", 512, 512, 3, 1, (15, 27), False),
    ("rn101 s4 expand", 512, 2048, 1, 1, (15, 27), True),
    ("rn18 l2 3x3", 128, 128, 3, 1, (60, 107), True),
    ("rn18 l3 3x3", 256, 256, 3, 1, (30, 54), True),
    ("rn18 l3 3x3 s2", 128, 256, 3, 2, (60, 107), False),
    ("rn18 l4 3x3", 512, 512, 3, 1, (15, 27), True),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between launches (the state inside a block)")
    ap.add_argument("--only", default="", help="substring of the shape names to run")
    ap.add_argument("--sels", default="1,2,3,4,0", help="kernel_select values to time")
    args = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    print("| conv | GFLOP (alg.) | general us | pair N=128 us | pair N=256 us | persistent pair us | auto us | auto alg. TFLOP/s |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    sels = [int(v) for v in args.sels.split(",")]
    for name, cin, cout, k, stride, hw, with_res in SHAPES:
        if args.only not in name:
            continue
        B = args.frames
        x = torch.randn(B, hw[0], hw[1], cin, generator=g).to(dev)
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        pc = ops.pack_conv_tc(w, torch.randn(cout, generator=g), device=dev, stride=stride)
        xs = ops.split_f16(x)
        Ho, Wo = (hw[0] - 1) // stride + 1, (hw[1] - 1) // stride + 1
        rs = ops.split_f16(torch.randn(B, Ho, Wo, cout, generator=g).to(dev)) if with_res else None
        gflop = 2.0 * B * Ho * Wo * cin * cout * k * k / 1e9
        row = []
        for sel in (1, 2, 3, 4, 0):
            if (sel == 3 and cout % 256) or sel not in sels:
                row.append(float("nan"))
                continue
            ts = []
            for it in range(args.reps + 3):
                if not args.no_flush:
                    flush.fill_(it & 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv2d_tc(xs, pc, res=rs, relu=True, out_f32=False, out_split=True, kernel_select=sel)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            row.append(ts[len(ts) // 2])
        print("| %s %d->%d k%d s%d %dx%d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.0f |" % (
            name, cin, cout, k, stride, hw[0], hw[1], gflop, row[0], row[1], row[2], row[3], row[4], gflop / row[4] * 1e3))


if __name__ == "__main__":
    main()
