"""-m gpu: the closed-form Gauss-Newton/CG path and the frame memory against the oracle (autograd restatement of
the reference) and against the golden fixtures produced by the executed reference."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fill_memory(prob):
    from frtm_vos_b200.model.memory import Memory
    from frtm_vos_b200 import ops
    cap = prob["samples"].shape[0]
    mem = Memory(cap, prob["samples"].shape[1:], prob["labels"].shape[1:], DEV, 0.1)
    mem.samples.copy_(prob["samples"]); mem.labels.copy_(prob["labels"])
    mem.pixel_weights.copy_(prob["pixel_weights"]); mem.weights.copy_(prob["weights"])
    st, uty = ops.build_stencil(mem.pixel_weights, mem.labels, mem.samples.shape[-2:])
    mem.stencil.copy_(st); mem.uty.copy_(uty)
    M = int((prob["weights"] > 0).sum())
    mem.state.copy_(torch.tensor([M, -1, -1, 0], dtype=torch.int32))
    return mem


def test_stencil_is_UtW2U():
    """S s == U^T (pw^2 * U s) and uty == U^T (pw^2 y), with U^T taken from autograd of F.interpolate."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(4)
    for (h, w), (H, W) in [((4, 7), (64, 112)), ((30, 54), (480, 854)), ((5, 5), (5, 5)), ((3, 4), (7, 9))]:
        K = 2
        pw = 0.5 + torch.rand(K, 1, H, W, generator=g)
        y = torch.rand(K, 1, H, W, generator=g)
        s = torch.randn(K, 1, h, w, generator=g, requires_grad=True)
        up = F.interpolate(s, (H, W), mode="bilinear", align_corners=False)
        ref_v, = torch.autograd.grad(up, s, pw * pw * up.detach())
        ref_t, = torch.autograd.grad(up, s, pw * pw * y)
        st, uty = ops.build_stencil(pw.to(DEV), y.to(DEV), (h, w))
        v = torch.empty(K, h, w, device=DEV)
        sw = torch.ones(K, device=DEV)
        lib().stencil_apply(ptr(st), ptr(s.detach().reshape(K, h, w).contiguous().to(DEV)), ptr(uty), ptr(sw), K, h, w, 0,
                            ptr(v), stream())
        scale = ref_v.abs().max().item()
        assert (v.cpu() - ref_v[:, 0]).abs().max() < 2e-6 * max(scale, 1.0), ((h, w), (H, W))
        assert (uty.cpu() - ref_t[:, 0]).abs().max() < 2e-6 * max(ref_t.abs().max().item(), 1.0)


def test_update_phase_matches_reference_golden(golden):
    """P2: two consecutive 10-iteration updates (persistent p / rho / r_prev) vs the executed reference."""
    from frtm_vos_b200.model.discriminator import DiscriminatorLoss
    from frtm_vos_b200.model.optimizer import GaussNewtonCG
    from frtm_vos_b200.lib.tensorlist import TensorList
    from frtm_vos_b200 import ops
    g = golden("update")
    prob = GI.update_problem()
    mem = _fill_memory(prob)
    filt = prob["F0"].clone().to(DEV)
    problem = DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-2,), precond=(1e-2,),
                                sample_weights=mem.weights, net=None, pixel_weighting=mem.pixel_weights, memory=mem)
    opt = GaussNewtonCG(problem, TensorList([filt]), fletcher_reeves=False, standard_alpha=True,
                        direction_forget_factor=(1 - 0.1) ** 750)
    opt.run((10,))
    scale = np.abs(g["F1"]).max()
    assert np.abs(filt.cpu().numpy() - g["F1"]).max() < 1e-5 * max(scale, 1.0)
    assert np.abs(opt.p[0].cpu().numpy() - g["p1"]).max() < 1e-3 * np.abs(g["p1"]).max()
    assert abs(float(opt.rho) - float(g["rho1"])) < 1e-3 * abs(float(g["rho1"]))
    # second run after a memory change
    s3, l3 = torch.from_numpy(g["s3"]).to(DEV), torch.from_numpy(g["l3"]).to(DEV)
    mem.samples[3] = s3
    mem.labels[3] = l3
    st, uty = ops.build_stencil(mem.pixel_weights[3:4], mem.labels[3:4], mem.samples.shape[-2:])
    mem.stencil[3] = st[0]; mem.uty[3] = uty[0]
    opt.run((10,))
    assert np.abs(filt.cpu().numpy() - g["F2"]).max() < 1e-5 * max(np.abs(g["F2"]).max(), 1.0)


@pytest.mark.parametrize("cap,M,c,h,w", [(16, 12, 96, 4, 7), (6, 5, 96, 30, 54), (4, 4, 64, 9, 13), (3, 3, 112, 17, 8)])
def test_tensor_core_operator_matches_cuda_core_operator(cap, M, c, h, w):
    """The tcgen05 operator kernel (split tile images) against the CUDA-core kernel (fp32 samples) through the C ABI:
    same filter after RHS + 5 CG iterations, and the phase-1 score map against a float64 conv2d."""
    from frtm_vos_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(c + h)
    X = torch.zeros(cap, c, h, w)
    X[:M] = torch.randn(M, c, h, w, generator=g) * 0.5
    S = (torch.rand(cap, 9, h, w, generator=g) * 4.0).to(DEV)
    T = torch.randn(cap, h, w, generator=g).to(DEV)
    sw = torch.zeros(cap)
    sw[:M] = torch.rand(M, generator=g) + 0.1
    sw = (sw / sw.sum()).to(DEV)
    f0 = torch.randn(c * 9, generator=g) * 0.05
    Xd = X.to(DEV)
    L = lib()
    nb = L.split_sample_bytes(c, h * w)
    assert nb == -(-h * w // 128) * 2 * 2 * c * 64 * 2 + -(-h * w // 256) * 10 * 256 * 4
    XS = torch.zeros(cap, nb, dtype=torch.uint8, device=DEV)
    L.split_samples(ptr(Xd), ptr(S), ptr(T), cap, c, h * w, ptr(XS), stream())
    nbytes = L.gn_update_workspace(cap, c, h, w)
    ws = torch.empty(nbytes // 4, device=DEV)
    arr = (ctypes.c_int * 1)(5)
    npad = (h + 2) * (w + 2)
    dbg = torch.zeros(2 * npad, device=DEV)
    res = {}
    for name, split in (("cuda_core", None), ("tensor_core", XS)):
        filt = f0.clone().to(DEV)
        st = torch.zeros(2 * c * 9 + 4, device=DEV)
        L.gn_debug_dump(ptr(dbg) if split is not None else None)
        try:
            L.gn_update(ptr(Xd), ptr(split), ptr(S), ptr(T), ptr(sw), cap, c, h, w, ptr(filt), ptr(st), arr, 1, 1e-2, 1e-2,
                        0.9 ** 750, None, 10, ptr(ws), nbytes, stream())
            torch.cuda.synchronize()
        finally:
            L.gn_debug_dump(None)
        res[name] = filt.cpu()
    ref = F.conv2d(X[0:1].double(), f0.double().view(1, c, 3, 3), padding=1)[0, 0].float()
    sp = dbg[:npad].view(h + 2, w + 2)[1:-1, 1:-1].cpu()
    assert (sp - ref).abs().max() < 4e-6 * max(ref.abs().max().item(), 1.0)
    step = (res["cuda_core"] - f0).abs().max().item()
    assert step > 1e-3                                         # the update did something
    # two different summation orders through 5 CG iterations: same budget as the golden parity test (1e-5)
    assert (res["tensor_core"] - res["cuda_core"]).abs().max() < 1e-5 * max(res["cuda_core"].abs().max().item(), 1.0)


def test_memory_keeps_split_image_in_step():
    """Memory.update (device-side slot choice) writes the operator image of the inserted sample."""
    from frtm_vos_b200.model.memory import Memory
    from frtm_vos_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(11)
    c, h, w, H, W = 96, 4, 7, 64, 112
    mem = Memory(6, (c, h, w), (1, H, W), DEV, 0.1)
    feats = torch.randn(5, c, h, w, generator=g).to(DEV)
    labels = (torch.rand(5, 1, H, W, generator=g) > 0.5).to(DEV)
    pw = (0.5 + torch.rand(5, 1, H, W, generator=g)).to(DEV)
    mem.initialize(feats, labels, pw)
    for k in range(3):
        f = torch.randn(1, c, h, w, generator=g).to(DEV)
        mem.update(f, torch.rand(1, 1, H, W, generator=g).to(DEV), (0.5 + torch.rand(1, 1, H, W, generator=g)).to(DEV))
    want = torch.zeros_like(mem.split)
    lib().split_samples(ptr(mem.samples), ptr(mem.stencil), ptr(mem.uty), 6, c, h * w, ptr(want), stream())
    assert torch.equal(want, mem.split)
    assert int((mem.split != 0).sum()) > 0


def test_update_gate_skips_on_device():
    from frtm_vos_b200.model.discriminator import DiscriminatorLoss
    from frtm_vos_b200.model.optimizer import GaussNewtonCG
    from frtm_vos_b200.lib.tensorlist import TensorList
    prob = GI.update_problem()
    mem = _fill_memory(prob)
    filt = prob["F0"].clone().to(DEV)
    problem = DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-2,), precond=(1e-2,),
                                sample_weights=mem.weights, net=None, pixel_weighting=mem.pixel_weights, memory=mem)
    opt = GaussNewtonCG(problem, TensorList([filt]), fletcher_reeves=False, standard_alpha=True,
                        direction_forget_factor=(1 - 0.1) ** 750)
    opt.run((5,), gate_count=torch.tensor([9], dtype=torch.int32, device=DEV), min_px=10)
    assert torch.equal(filt.cpu(), prob["F0"]) and opt.p is None
    opt.run((5,), gate_count=torch.tensor([10], dtype=torch.int32, device=DEV), min_px=10)
    assert not torch.equal(filt.cpu(), prob["F0"])


def test_init_problem_teacher_forced(golden):
    """P3: RHS and one J^T J product of the joint project/filter problem at a fixed point vs the executed reference."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200._lib import lib, ptr, stream
    g = golden("init_step")
    ip = GI.init_problem()
    K, C, h, w = ip["x"].shape
    x_nhwc = ip["x"].permute(0, 2, 3, 1).contiguous().to(DEV)
    y = ip["y"].to(DEV)
    pw = ops.pixel_weights(y, 0.1, False)
    assert np.abs(pw.cpu().numpy() - g["pw"]).max() < 1e-6
    st, uty = ops.build_stencil(pw, y, (h, w))
    sw = torch.tensor([2.0, 1, 1, 1, 1]) / 6.0
    P, Fw = ip["P0"].clone().to(DEV), ip["F0"].clone().to(DEV)
    dP, dF = ip["dP"].to(DEV), ip["dF"].to(DEV)
    outs = [torch.empty_like(P), torch.empty_like(Fw), torch.empty_like(P), torch.empty_like(Fw)]
    L = lib()
    nbytes = L.gn_init_workspace(K, C, 96, h, w)
    ws = torch.empty(nbytes // 4, device=DEV)
    L.gn_init_probe(ptr(x_nhwc), ptr(st), ptr(uty), ptr(sw.to(DEV)), K, C, 96, h, w, ptr(P), ptr(Fw), ptr(dP), ptr(dF), 1e-4,
                    1e-2, ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ptr(ws), nbytes, stream())
    for got, key in zip(outs, ("bP", "bF", "AP", "AF")):
        ref = g[key]
        assert np.abs(got.cpu().numpy() - ref).max() < 1e-5 * np.abs(ref).max(), key


def test_init_free_running_functional(golden):
    """Free-running init diverges at ulp level by nature (SURVEY finding 8): compare the fitted score map."""
    from frtm_vos_b200.model.discriminator import Discriminator
    g = golden("init_step")
    ip = GI.init_problem()
    d = Discriminator(in_channels=ip["x"].shape[1], init_iters=(5, 10), update_iters=(5,), memory_size=8,
                      CG_forgetting_rate=750, pixel_weighting=dict(method="hinge", tf=0.1), device=DEV)
    d.project.weight.data.copy_(ip["P0"]); d.filter.weight.data.copy_(ip["F0"])
    d.init(ip["x"].to(DEV), ip["y"].byte().to(DEV))
    s = d(ip["x"].to(DEV)).cpu().numpy()
    assert np.isfinite(s).all()
    assert np.abs(s - g["s_fin"]).max() < 5e-2
    assert np.allclose(d.memory.weights.cpu().numpy(), g["w_fin"], atol=1e-7)
    assert d.memory.current_size == 5


def test_memory_trace_golden(golden):
    from frtm_vos_b200.model.memory import Memory
    g = golden("memory")
    z = torch.zeros(5, 1, 1, 1, device=DEV)
    m = Memory(10, (1, 1, 1), (1, 1, 1), DEV, 0.1)
    m.initialize(z, z, z)
    assert np.allclose(m.weights.cpu().numpy(), g["weights"][0], atol=1e-7)
    for i in range(25):
        m.update(z[0], z[0], z[0])
        assert m.previous_replace_ind == int(g["replace"][i])
        assert np.allclose(m.weights.cpu().numpy(), g["weights"][i + 1], atol=2e-7)
    # gated insert leaves everything untouched
    before = m.weights.clone()
    m.update(z[0], z[0], z[0], gate_count=torch.tensor([3], dtype=torch.int32, device=DEV))
    assert torch.equal(before, m.weights) and int(m.state[2]) == -1
