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


def _operator_problem(cap, M, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.zeros(cap, c, h, w)
    act = torch.randperm(cap, generator=g)[:M]                 # active slots anywhere in the buffer, as after replacements
    X[act] = torch.randn(M, c, h, w, generator=g) * 0.5
    S = torch.rand(cap, 9, h, w, generator=g) * 4.0
    T = torch.randn(cap, h, w, generator=g)
    sw = torch.zeros(cap)
    sw[act] = torch.rand(M, generator=g) + 0.1
    sw = sw / sw.sum()
    f0 = torch.randn(c * 9, generator=g) * 0.05
    return X, S, T, sw, f0


def _operator_fp64(X, S, T, sw, p, use_y):
    """sum_i X_i^T [ sw_i (S_i (X_i * p) - use_y t_i) ] in float64 (stencil form, SURVEY.md §8(d) form S)."""
    cap, c, h, w = X.shape
    Xd, pd = X.double(), p.double().view(1, c, 3, 3)
    s = F.conv2d(Xd, pd, padding=1)[:, 0]                                            # (cap,h,w)
    sp = F.pad(s, (1, 1, 1, 1))
    v = torch.zeros_like(s)
    for t in range(9):
        v += S[:, t].double() * sp[:, t // 3:t // 3 + h, t % 3:t % 3 + w]
    if use_y:
        v -= T.double()
    v = v * sw.double().view(-1, 1, 1)
    # gradient of sum(conv(X, p) * v) w.r.t. p
    g = torch.nn.grad.conv2d_weight(Xd.reshape(1, cap * c, h, w), (cap, c, 3, 3), v.view(1, cap, h, w), padding=1, groups=cap)
    return g.view(cap, c, 3, 3).sum(0).reshape(-1)


def _cg_fp64(X, S, T, sw, f0, n_cg, reg=1e-2, precond=1e-2):
    """RHS + n_cg Polak-Ribiere iterations exactly as frtm_gn_update runs them (fresh state), in float64."""
    f = f0.double().clone()
    b = -(_operator_fp64(X, S, T, sw, f, True) + reg * reg * f)
    r, x, p, rprev, rho = b.clone(), torch.zeros_like(b), None, None, 1.0
    for it in range(n_cg):
        z = r / precond
        rho1, rho = rho, float(r @ z)
        if p is None:
            p = z.clone()
        else:
            beta = max((rho - float(rprev @ z)) / rho1, 0.0)
            p = z + p * beta
        q = _operator_fp64(X, S, T, sw, p, False) + reg * reg * p
        alpha = rho / float(p @ q)
        rprev = r.clone()
        x = x + alpha * p
        if it < n_cg - 1:
            r = r - alpha * q
    return (f + x).float()


# (cap, M, c, h, w, CG iterations): the last shape has 18 pixels for 864 unknowns — one iteration, or CG amplifies rounding
OPERATOR_SHAPES = [(16, 12, 96, 4, 7, 5), (6, 5, 96, 30, 54, 5), (4, 4, 64, 9, 13, 5), (3, 3, 112, 17, 8, 5), (5, 4, 96, 8, 14, 5),
                   (3, 3, 96, 45, 80, 5), (4, 3, 96, 11, 84, 5), (3, 2, 96, 2, 9, 1)]


@pytest.mark.parametrize("cap,M,c,h,w,n_cg", OPERATOR_SHAPES)
def test_operator_kernels_match_fp64_and_each_other(cap, M, c, h, w, n_cg):
    """Every operator kernel that takes the shape (cluster-resident, single-pass sliding window, two-pass tcgen05, CUDA
    cores; selected per call through ``operator_select`` of the C ABI) against a float64 evaluation of the same RHS + CG
    iterations."""
    from frtm_vos_b200._lib import lib, ptr, stream
    X, S, T, sw, f0 = _operator_problem(cap, M, c, h, w, seed=c + h)
    ref = _cg_fp64(X, S, T, sw, f0, n_cg)
    Xd, Sd, Td, swd = X.to(DEV), S.to(DEV), T.to(DEV), sw.to(DEV)
    L = lib()
    nb = L.split_sample_bytes(c, h * w)
    assert nb == -(-h * w // 128) * 2 * 2 * c * 64 * 2 + -(-h * w // 256) * 10 * 256 * 4
    XS = torch.zeros(cap, nb, dtype=torch.uint8, device=DEV)
    L.split_samples(ptr(Xd), ptr(Sd), ptr(Td), cap, c, h * w, ptr(XS), stream())
    nbytes = L.gn_update_workspace(cap, c, h, w)
    ws = torch.empty(nbytes // 4, device=DEV)
    arr = (ctypes.c_int * 1)(n_cg)
    step = (ref - f0).abs().max().item()
    assert step > 1e-3                                         # the update did something
    scale = max(ref.abs().max().item(), 1.0)
    ran = []
    for sel, name in ((4, "cluster"), (3, "single-pass"), (2, "two-pass"), (1, "cuda-core"), (0, "auto")):
        filt = f0.clone().to(DEV)
        st = torch.zeros(2 * c * 9 + 4, device=DEV)
        try:
            L.gn_update(ptr(Xd), ptr(XS), ptr(Sd), ptr(Td), ptr(swd), cap, c, h, w, ptr(filt), ptr(st), arr, 1, 1e-2, 1e-2,
                        0.9 ** 750, None, 10, sel, ptr(ws), nbytes, stream())
        except RuntimeError as e:
            assert "not supported" in str(e) and sel in (2, 3, 4), (name, str(e))
            continue
        torch.cuda.synchronize()
        ran.append(sel)
        # two summation orders through 5 CG iterations: the budget of the golden parity test (1e-5)
        assert (filt.cpu() - ref).abs().max().item() < 1e-5 * scale, (name, (filt.cpu() - ref).abs().max().item())
    assert 1 in ran and 0 in ran
    if c == 96 and 8 <= w <= 84:
        assert 3 in ran and 4 in ran                           # the production shapes run the single-pass kernels
    assert L.gn_operator_kind(c, h, w) == (3 if 3 in ran else 2 if 2 in ran else 1)


@pytest.mark.parametrize("cap,M,h,w,H,W,n_cg", [(80, 72, 30, 54, 480, 854, 10), (32, 32, 45, 80, 720, 1280, 10)])
def test_update_at_baseline_shapes_matches_oracle_autograd(cap, M, h, w, H, W, n_cg):
    """The filter update at the BASELINE configurations (config 3: cap 80 at 480p with the memory nearly full and a slot
    being replaced; config 5: cap 32 full at 720p) against the oracle's autograd GaussNewtonCG (the reference's own
    double-backward operator, model/optimizer.py:77-157) over two consecutive runs with a replacement in between."""
    from oracle import frtm_ref as R
    from frtm_vos_b200.model.memory import Memory
    from frtm_vos_b200.model.discriminator import DiscriminatorLoss
    from frtm_vos_b200.model.optimizer import GaussNewtonCG
    from frtm_vos_b200.lib.tensorlist import TensorList
    from frtm_vos_b200 import ops
    c = 96
    g = torch.Generator().manual_seed(1000 + h)
    act = torch.randperm(cap, generator=g)[:M]
    samples = torch.zeros(cap, c, h, w)
    samples[act] = torch.randn(M, c, h, w, generator=g) * 0.3
    # soft masks: a blurred blob per sample, like merged segmentation probabilities
    lab = torch.zeros(cap, 1, H, W)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    for k in act.tolist():
        cy, cx = float(torch.rand(1, generator=g)) * H, float(torch.rand(1, generator=g)) * W
        ry, rx = H * (0.1 + 0.2 * float(torch.rand(1, generator=g))), W * (0.1 + 0.2 * float(torch.rand(1, generator=g)))
        lab[k, 0] = torch.sigmoid(4.0 * (1.0 - ((yy - cy) / ry) ** 2 - ((xx - cx) / rx) ** 2))
    wts = torch.zeros(cap)
    wts[act] = torch.rand(M, generator=g) + 0.05
    wts = wts / wts.sum()
    F0 = torch.randn(1, c, 3, 3, generator=g) * 0.02
    mem = Memory(cap, (c, h, w), (1, H, W), DEV, 0.1)
    mem.samples.copy_(samples); mem.labels.copy_(lab); mem.weights.copy_(wts)
    pw = ops.pixel_weights(mem.labels, 0.1, True)
    mem.pixel_weights.copy_(pw)
    st, uty = ops.build_stencil(mem.pixel_weights, mem.labels, (h, w))
    mem.stencil.copy_(st); mem.uty.copy_(uty)
    mem.state.copy_(torch.tensor([M, -1, -1, 0], dtype=torch.int32))
    filt = F0.clone().to(DEV)
    problem = DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-2,), precond=(1e-2,), sample_weights=mem.weights,
                                net=None, pixel_weighting=mem.pixel_weights, memory=mem)
    opt = GaussNewtonCG(problem, TensorList([filt]), fletcher_reeves=False, standard_alpha=True,
                        direction_forget_factor=0.9 ** 750)
    om = R.FrameMemory(cap, (c, h, w), (1, H, W), "cpu", 0.1)
    om.samples.copy_(samples); om.labels.copy_(lab); om.pixel_weights.copy_(pw.cpu()); om.weights.copy_(wts)
    Fo = F0.clone()
    oopt = R.GaussNewtonCGRef(R.GNProblem(om, (1e-2,), (1e-2,), False), [Fo], 0.9 ** 750)
    opt.run((n_cg,))
    oopt.run((n_cg,))
    scale = max(Fo.abs().max().item(), 1.0)
    assert (Fo - F0).abs().max().item() > 1e-3
    err1 = (filt.cpu() - Fo).abs().max().item()
    assert err1 < 1e-5 * scale, err1
    # replace the sample of one slot (as Memory.update does on a full memory) and run again with the persistent CG state
    r = int(act[0])
    new_s = torch.randn(c, h, w, generator=g) * 0.3
    mem.samples[r] = new_s.to(DEV); om.samples[r] = new_s
    opt.run((n_cg,))
    oopt.run((n_cg,))
    err2 = (filt.cpu() - Fo).abs().max().item()
    assert err2 < 1e-5 * scale, err2


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


def test_memory_insert_block_equals_sequential_inserts():
    """frtm_memory_insert_block (all inserts of a track block in two launches) against the same inserts applied frame by
    frame with Memory.update (model/memory.py:59-92): two objects whose memories fill up and start replacing inside the
    block, some (frame, object) pairs gated out — every buffer, the sample weights, the policy state and the operator
    images bit-identical; the slots reported by the block call are the replace indices of the sequential run."""
    import ctypes
    from frtm_vos_b200.model.memory import Memory
    from frtm_vos_b200._lib import lib, ptr, stream
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(5)
    c, h, w, H, W, cap, n, nF = 96, 4, 7, 64, 112, 8, 2, 8

    def fresh():
        mems = []
        gi = torch.Generator().manual_seed(9)
        for o in range(n):
            m = Memory(cap, (c, h, w), (1, H, W), DEV, 0.1)
            K = 5
            m.initialize(torch.randn(K, c, h, w, generator=gi).to(DEV), (torch.rand(K, 1, H, W, generator=gi) > 0.5).to(DEV),
                         (0.5 + torch.rand(K, 1, H, W, generator=gi)).to(DEV))
            mems.append(m)
        return mems

    feats = torch.randn(nF * n, c, h, w, generator=g).to(DEV)
    ys = torch.rand(nF * n, 1, H, W, generator=g).to(DEV)
    pw = (0.5 + torch.rand(nF * n, 1, H, W, generator=g)).to(DEV)
    st, uty = ops.build_stencil(pw, ys, (h, w))
    counts = torch.full((nF, n), 100, dtype=torch.int32)
    counts[2, 0] = 3
    counts[5, 1] = 0                                        # below min_px = 10: skipped on the device
    counts = counts.to(DEV)

    seq = fresh()
    slots_seq = []
    for f in range(nF):
        for o in range(n):
            j = f * n + o
            seq[o].update(feats[j:j + 1], ys[j:j + 1], pw[j:j + 1], st[j:j + 1], uty[j:j + 1], gate_count=counts[f, o:o + 1], min_px=10)
            slots_seq.append(int(seq[o].state[2]))

    blk = fresh()
    rows = [[m.samples.data_ptr() for m in blk], [m.labels.data_ptr() for m in blk], [m.pixel_weights.data_ptr() for m in blk],
            [m.stencil.data_ptr() for m in blk], [m.uty.data_ptr() for m in blk], [m.split.data_ptr() for m in blk],
            [m.weights.data_ptr() for m in blk], [m.state.data_ptr() for m in blk]]
    table = torch.tensor([v for r in rows for v in r], dtype=torch.int64).to(DEV)
    slots = torch.empty(nF * n, dtype=torch.int32, device=DEV)
    lib().memory_insert_block(ptr(table), n, nF, cap, 0.1, ptr(counts), 10, ptr(feats), feats[0].numel(), ptr(ys), ptr(pw), H * W,
                              ptr(st), ptr(uty), h * w, 1, 1, ptr(slots), stream())
    assert slots.cpu().tolist() == slots_seq
    used = [s for s in slots_seq[0::n] if s >= 0]
    assert slots_seq[2 * n + 0] == -1 and slots_seq[5 * n + 1] == -1
    assert len(set(used)) >= 3 and len(used) > len(set(used))       # the full memory replaces a slot twice inside the block
    for a, b in zip(seq, blk):
        for name in ("samples", "labels", "pixel_weights", "stencil", "uty", "weights", "state", "split"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
