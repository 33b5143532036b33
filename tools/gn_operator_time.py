"""Times one batched filter update (RHS + n_cg operator applications) per operator kernel on synthetic memories:

    python tools/gn_operator_time.py [n_obj] [M] [cap] [h] [w] [n_cg] [kernels]

defaults: 3 objects x 69 of 80 samples at 30x54, 5 CG iterations (the config-2 update), kernels = "4,3,2,1"
(4 = cluster-resident, 3 = single-pass sliding window, 2 = two-pass tcgen05, 1 = CUDA cores).  Also checks the filters of the kernels against each other.
Prints algorithmic GB/s (form S of SURVEY.md §8(d)) and the fraction of the measured HBM peak.
"""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frtm_vos_b200._lib import lib, ptr, stream  # noqa: E402
DEV = "cuda:0"
arg = lambda k, d: type(d)(sys.argv[k]) if len(sys.argv) > k else d
n_obj, M, cap, h, w, n_cg = arg(1, 3), arg(2, 69), arg(3, 80), arg(4, 30), arg(5, 54), arg(6, 5)
kernels = [int(v) for v in arg(7, "4,3,2,1").split(",")]
c = 96
peak = 6533.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]
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
    f0 = (torch.randn(c * 9, generator=g) * 0.05).to(DEV)
    gate = torch.tensor([100], dtype=torch.int32, device=DEV)
    objs.append(dict(X=X, S=S, T=T, sw=sw, XS=XS, f0=f0, filt=f0.clone(), st=torch.zeros(2 * c * 9 + 4, device=DEV), gate=gate))
rows = [[o[k].data_ptr() for o in objs] for k in ("X", "S", "T", "sw", "filt", "st", "gate", "XS")]
table = torch.tensor([v for r in rows for v in r], dtype=torch.int64).to(DEV)
nbytes = n_obj * L.gn_update_workspace(cap, c, h, w)
ws = torch.empty(nbytes // 4, device=DEV)
arr = (ctypes.c_int * 1)(n_cg)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
names = {1: "cuda-core", 2: "two-pass tcgen05", 3: "single-pass mma", 4: "cluster mma"}
results = {}


def run(sel):
    L.gn_update_batched(ptr(table), n_obj, 1, cap, c, h, w, arr, 1, 1e-2, 1e-2, 0.9 ** 750, 10, sel, ptr(ws), nbytes, stream())


for sel in kernels:
    for o in objs:
        o["filt"].copy_(o["f0"]); o["st"].zero_()
    try:
        run(sel)
    except RuntimeError as e:
        print("%s: not available for this shape (%s)" % (names[sel], e))
        continue
    torch.cuda.synchronize()
    results[sel] = torch.stack([o["filt"].clone() for o in objs]).cpu()
    for rep in range(3):
        run(sel)
    torch.cuda.synchronize()
    ms_all = []
    for rep in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(sel); e1.record(); torch.cuda.synchronize()
        ms_all.append(e0.elapsed_time(e1))
    ms_all.sort()
    ms = ms_all[len(ms_all) // 2]
    byt = n_obj * M * 4 * (c * h * w + 10 * h * w) + n_cg * n_obj * M * 4 * (c * h * w + 9 * h * w)
    print("%-18s %d objects x %d/%d samples %dx%d: %.3f ms per update (RHS + %d A.p; min %.3f) -> %.0f GB/s = %.3f of %.0f"
          % (names[sel], n_obj, M, cap, h, w, ms, n_cg, ms_all[0], byt / ms / 1e6, byt / ms / 1e6 / peak, peak))
base = results.get(1, None)
for sel, f in results.items():
    if base is not None and sel != 1:
        print("%s vs cuda-core: max |dF| = %.3e (|F| max %.3f, step %.3e)" % (
            names[sel], (f - base).abs().max().item(), base.abs().max().item(),
            (base - torch.stack([o["f0"] for o in objs]).cpu()).abs().max().item()))
