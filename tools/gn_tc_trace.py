"""Timeline of one CTA of the tensor-core GN/CG operator kernel (library must be built with -DGC_TRACE).
python tools/gn_tc_trace.py [n_obj]"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frtm_vos_b200._lib import lib, ptr, stream  # noqa: E402
DEV = "cuda:0"
n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M, cap, c, h, w = 69, 80, 96, 30, 54
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
    objs.append(dict(X=X, S=S, T=T, sw=sw, XS=XS, filt=(torch.randn(c * 9, generator=g) * 0.05).to(DEV),
                     st=torch.zeros(2 * c * 9 + 4, device=DEV), gate=torch.tensor([100], dtype=torch.int32, device=DEV)))
rows = [[o[k].data_ptr() for o in objs] for k in ("X", "S", "T", "sw", "filt", "st", "gate", "XS")]
table = torch.tensor([v for r in rows for v in r], dtype=torch.int64).to(DEV)
nbytes = n_obj * L.gn_update_workspace(cap, c, h, w)
ws = torch.empty(nbytes // 4, device=DEV)
arr = (ctypes.c_int * 1)(2)
npad = (h + 2) * (w + 2)
dbg = torch.zeros(2 * npad + 2 + 2 * 64 * 4 + 16, device=DEV)
for rep in range(3):
    L.gn_update_batched(ptr(table), n_obj, 1, cap, c, h, w, arr, 1, 1e-2, 1e-2, 0.9 ** 750, 10, ptr(ws), nbytes, stream())
torch.cuda.synchronize()
L.gn_debug_dump(ptr(dbg))
L.gn_update_batched(ptr(table), n_obj, 1, cap, c, h, w, arr, 1, 1e-2, 1e-2, 0.9 ** 750, 10, ptr(ws), nbytes, stream())
torch.cuda.synchronize()
L.gn_debug_dump(None)
raw = dbg[2 * npad + 2: 2 * npad + 2 + 512].cpu().numpy().view(np.uint64).reshape(4, 64)
t0 = raw[0, 0]
names = ["cta", "producer issue", "mma ready", "drain"]
for r in range(4):
    vals = [(k, (int(v) - int(t0)) / 1e3) for k, v in enumerate(raw[r]) if v != 0]
    print(names[r], " ".join("%d:%.1f" % kv for kv in vals))
