"""One warm batched filter update (3 objects x 69 of 80 samples, 30x54) for ncu launch lists: python tools/gn_tc_time.py [n_obj] [M]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frtm_vos_b200._lib import lib, ptr, stream  # noqa: E402
DEV = "cuda:0"
n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M = int(sys.argv[2]) if len(sys.argv) > 2 else 69
cap, c, h, w = 80, 96, 30, 54
L = lib()
g = torch.Generator().manual_seed(0)
objs = []
for o in range(n_obj):
    X = torch.zeros(cap, c, h, w); X[:M] = torch.randn(M, c, h, w, generator=g) * 0.5
    S = torch.rand(cap, 9, h, w, generator=g) * 4.0
    T = torch.randn(cap, h, w, generator=g)
    sw = torch.zeros(cap); sw[:M] = torch.rand(M, generator=g) + 0.1; sw /= sw.sum()
    X, S, T, sw = X.to(DEV), S.to(DEV), T.to(DEV), sw.to(DEV)
    XS = torch.zeros(cap, L.split_sample_bytes(c, h * w) // 2, dtype=torch.float16, device=DEV)
    L.split_samples(ptr(X), ptr(S), ptr(T), cap, c, h * w, ptr(XS), stream())
    filt = (torch.randn(c * 9, generator=g) * 0.05).to(DEV)
    st = torch.zeros(2 * c * 9 + 4, device=DEV)
    gate = torch.tensor([100], dtype=torch.int32, device=DEV)
    objs.append(dict(X=X, S=S, T=T, sw=sw, XS=XS, filt=filt, st=st, gate=gate))
rows = [[o[k].data_ptr() for o in objs] for k in ("X", "S", "T", "sw", "filt", "st", "gate", "XS")]
flat = [v for r in rows for v in r]
table = torch.tensor(flat, dtype=torch.int64).to(DEV)
nbytes = n_obj * L.gn_update_workspace(cap, c, h, w)
ws = torch.empty(nbytes // 4, device=DEV)
arr = (ctypes.c_int * 1)(5)
for has_split in (1, 0):
    for rep in range(3):
        L.gn_update_batched(ptr(table), n_obj, has_split, cap, c, h, w, arr, 1, 1e-2, 1e-2, 0.9 ** 750, 10, ptr(ws), nbytes, stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(10):
        L.gn_update_batched(ptr(table), n_obj, has_split, cap, c, h, w, arr, 1, 1e-2, 1e-2, 0.9 ** 750, 10, ptr(ws), nbytes, stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byt = n_obj * M * 4 * (c * h * w + 9 * h * w + h * w) + 5 * n_obj * M * 4 * (c * h * w + 9 * h * w)
    print("%s: %d objects x %d samples: %.3f ms per update (RHS + 5 A.p) -> %.0f GB/s" % ("tensor-core" if has_split else "cuda-core", n_obj, M, ms, byt / ms / 1e6))
