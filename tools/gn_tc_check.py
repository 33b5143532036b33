"""Tensor-core GN/CG operator kernel vs the CUDA-core kernel and vs torch (debug / timing aid, GPU only).
python tools/gn_tc_check.py"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frtm_vos_b200._lib import lib, ptr, stream  # noqa: E402

DEV = "cuda:0"


def run(cap, M, c, h, w, iters, seed=0, time_it=False):
    g = torch.Generator().manual_seed(seed)
    X = torch.zeros(cap, c, h, w)
    X[:M] = torch.randn(M, c, h, w, generator=g) * 0.5
    S = torch.rand(cap, 9, h, w, generator=g) * 4.0
    T = torch.randn(cap, h, w, generator=g)
    sw = torch.zeros(cap)
    sw[:M] = torch.rand(M, generator=g) + 0.1
    sw /= sw.sum()
    f0 = torch.randn(c * 9, generator=g) * 0.05
    X, S, T, sw = X.to(DEV), S.to(DEV), T.to(DEV), sw.to(DEV)
    L = lib()
    nb = L.split_sample_bytes(c, h * w)
    XS = torch.zeros(cap, nb // 2, dtype=torch.float16, device=DEV)
    L.split_samples(ptr(X), ptr(S), ptr(T), cap, c, h * w, ptr(XS), stream())
    nbytes = L.gn_update_workspace(cap, c, h, w)
    ws = torch.empty(nbytes // 4, device=DEV)
    arr = (ctypes.c_int * 1)(iters)
    npad = (h + 2) * (w + 2)
    out = {}
    for name, split in (("simt", None), ("tc", XS)):
        filt = f0.clone().to(DEV)
        st = torch.zeros(2 * c * 9 + 4, device=DEV)
        dbg = torch.zeros(2 * npad, device=DEV)
        L.gn_debug_dump(ptr(dbg) if split is not None else None)
        L.gn_update(ptr(X), ptr(split), ptr(S), ptr(T), ptr(sw), cap, c, h, w, ptr(filt), ptr(st), arr, 1, 1e-2, 1e-2,
                    0.9 ** 750, None, 10, ptr(ws), nbytes, stream())
        torch.cuda.synchronize()
        L.gn_debug_dump(None)
        out[name] = (filt.cpu(), st.cpu(), dbg.cpu(), ws[:cap * c * 9].clone().cpu())
    # the dump holds the RHS launch of sample 0: padded scores of X[0] with f0, then padded v
    npd = (h + 2) * (w + 2)
    sp = out["tc"][2][:npd].view(h + 2, w + 2)[1:-1, 1:-1]
    ref = F.conv2d(X[0:1].cpu().double(), f0.double().view(1, c, 3, 3), padding=1)[0, 0].float()
    print("   scores of sample 0: |s| max %.3e, diff vs conv2d %.3e" % (ref.abs().max(), (sp - ref).abs().max()))
    fs, ft = out["simt"][0], out["tc"][0]
    print("shape cap=%d M=%d c=%d %dx%d iters=%d : |F_tc - F_simt| max %.3e (|F| max %.3e, |dF| max %.3e)" % (
        cap, M, c, h, w, iters, (fs - ft).abs().max(), fs.abs().max(), (fs - f0).abs().max()))
    ps, pt = out["simt"][3], out["tc"][3]
    print("   last partials: max diff %.3e of max %.3e" % ((ps - pt).abs().max(), ps.abs().max()))
    if time_it:
        for name, split in (("simt", None), ("tc", XS)):
            filt = f0.clone().to(DEV)
            st = torch.zeros(2 * c * 9 + 4, device=DEV)
            arr10 = (ctypes.c_int * 1)(10)
            for _ in range(3):
                L.gn_update(ptr(X), ptr(split), ptr(S), ptr(T), ptr(sw), cap, c, h, w, ptr(filt), ptr(st), arr10, 1, 1e-2, 1e-2,
                            0.9 ** 750, None, 10, ptr(ws), nbytes, stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                L.gn_update(ptr(X), ptr(split), ptr(S), ptr(T), ptr(sw), cap, c, h, w, ptr(filt), ptr(st), arr10, 1, 1e-2, 1e-2,
                            0.9 ** 750, None, 10, ptr(ws), nbytes, stream())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            byt = M * 4 * (c * h * w + 9 * h * w) * 11
            print("   %s: %.3f ms per update (RHS + 10 A.p) -> %.0f GB/s algorithmic" % (name, ms, byt / ms / 1e6))


if __name__ == "__main__":
    run(16, 12, 96, 4, 7, 5)
    run(16, 12, 96, 30, 54, 5)
    run(8, 8, 96, 45, 80, 3)
    run(80, 69, 96, 30, 54, 10, time_it=True)
    run(32, 32, 96, 45, 80, 10, time_it=True)
    run(8, 5, 64, 9, 13, 2)
