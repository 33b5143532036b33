"""-m gpu: every C-ABI kernel against the same operation in plain PyTorch fp32 on the CPU (the ops the reference
calls), on seeded inputs, incl. ragged / edge shapes.  Tolerances are written per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from frtm_vos_b200 import ops
    return ops


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw,B", [
    (64, 64, 3, 1, 1, (30, 54), 2),       # plain 3x3
    (256, 96, 1, 1, 0, (30, 54), 1),      # projection-like 1x1
    (128, 256, 3, 2, 1, (61, 107), 1),    # stride 2, odd size
    (64, 128, 1, 2, 0, (60, 107), 1),     # downsample 1x1 s2
    (3, 64, 7, 2, 3, (97, 131), 2),       # stem, 3 input channels (padded to 4)
    (65, 65, 3, 1, 1, (15, 27), 3),       # 65 -> 65 TSE conv (padded to 68)
    (65, 64, 3, 1, 1, (15, 27), 1),
    (64, 32, 3, 1, 1, (40, 44), 1),       # BN=32 tile variant
    (512, 2048, 1, 1, 0, (8, 14), 1),
    (64, 64, 3, 1, 1, (1, 1), 1),         # degenerate 1x1 map
])
def test_conv2d(cin, cout, k, stride, pad, hw, B):
    ops = _ops()
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(B, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, stride, pad)
    res = torch.randn(ref.shape, generator=g)
    ref_full = F.relu(ref + res)
    cpad = (cin + 3) // 4 * 4
    xp = torch.zeros(B, cpad, *hw)
    xp[:, :cin] = x
    pc = ops.pack_conv(w, b, stride=stride, pad=pad, device=DEV, cin_pad=cpad)
    y, y_nchw = ops.conv2d(_nhwc(xp).to(DEV), pc, nchw=True)
    tol = 2e-5 * max(1.0, ref.abs().max().item())
    assert (_nchw(y.cpu()) - ref).abs().max() < tol
    assert (y_nchw.cpu() - ref).abs().max() < tol
    y2 = ops.conv2d(_nhwc(xp).to(DEV), pc, res=_nhwc(res).to(DEV), relu=True)
    assert (_nchw(y2.cpu()) - ref_full).abs().max() < tol


def test_conv2d_into_wide_buffer():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 64, 9, 11, generator=g)
    w = torch.randn(64, 64, 1, 1, generator=g) / 8
    pc = ops.pack_conv(w, None, device=DEV)
    out = torch.full((1, 9, 11, 68), 7.0, device=DEV)
    ops.conv2d(_nhwc(x).to(DEV), pc, out=out, coff=0)
    ref = F.conv2d(x, w)
    assert (_nchw(out[..., :64].cpu()) - ref).abs().max() < 1e-5
    assert torch.all(out[..., 64:] == 7.0)          # neighbours untouched


def test_bad_arguments_raise():
    ops = _ops()
    w = torch.randn(8, 6, 1, 1)
    pc = ops.pack_conv(w, None, device=DEV, cin_pad=6)
    with pytest.raises(RuntimeError, match="multiples of 4"):
        ops.conv2d(torch.zeros(1, 4, 4, 6, device=DEV), pc)


def test_normalize_and_maxpool():
    ops = _ops()
    from oracle import frtm_ref as R
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (2, 3, 37, 53), generator=g, dtype=torch.uint8)
    y = ops.normalize_u8(img.to(DEV)).cpu()
    assert torch.equal(_nchw(y)[:, :3], R.normalize_image(img))            # bit-exact
    assert torch.all(y[..., 3] == 0)
    x = torch.randn(2, 64, 37, 53, generator=g)
    p, p_nchw = ops.maxpool3x3s2(_nhwc(x).to(DEV), nchw=True)
    ref = F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(_nchw(p.cpu()), ref) and torch.equal(p_nchw.cpu(), ref)


@pytest.mark.parametrize("src,dst,C", [((30, 54), (480, 854), 1), ((15, 27), (30, 54), 64), ((480, 856), (480, 854), 32),
                                         ((30, 54), (15, 27), 1), ((4, 7), (64, 112), 1), ((1, 1), (5, 9), 64)])
def test_resize_bilinear(src, dst, C):
    ops = _ops()
    g = torch.Generator().manual_seed(src[0] * 7 + C)
    x = torch.randn(2, C, *src, generator=g)
    ref = F.interpolate(x, dst, mode="bilinear", align_corners=False)
    y = ops.resize_bilinear(_nhwc(x).to(DEV), dst)
    assert (_nchw(y.cpu()) - ref).abs().max() < 2e-6


def test_pyrup_bicubic():
    ops = _ops()
    from oracle import frtm_ref as R
    g = torch.Generator().manual_seed(5)
    for shp in [(2, 64, 13, 17), (1, 32, 1, 1), (1, 32, 2, 3)]:
        x = torch.randn(*shp, generator=g)
        ref = R.pyr_up_bicubic(x)
        y = ops.pyrup_bicubic(_nhwc(x).to(DEV))
        assert (_nchw(y.cpu()) - ref).abs().max() < 2e-6


def test_gap_and_cab():
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 64, 30, 54, generator=g)
    pooled = ops.global_avgpool(_nhwc(x).to(DEV))
    assert (pooled.cpu() - x.mean(dim=(2, 3))).abs().max() < 1e-6
    deeper = torch.randn(3, 64, 30, 54, generator=g)
    dp = deeper.mean(dim=(2, 3))
    w1, b1 = torch.randn(64, 128, generator=g) / 11, torch.randn(64, generator=g)
    w2, b2 = torch.randn(64, 64, generator=g) / 8, torch.randn(64, generator=g)
    gate = torch.sigmoid(F.linear(F.relu(F.linear(torch.cat((x.mean(dim=(2, 3)), dp), 1), w1, b1)), w2, b2))
    ref = x * gate[:, :, None, None] + deeper
    out = ops.cab(_nhwc(x).to(DEV), pooled, dp.to(DEV), _nhwc(deeper).to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    assert (_nchw(out.cpu()) - ref).abs().max() < 1e-5
    low = torch.randn(3, 64, 15, 27, generator=g)                       # deeper level at its own resolution: resized in the kernel
    out_r = ops.cab(_nhwc(x).to(DEV), pooled, dp.to(DEV), _nhwc(low).to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    ref_r = x * gate[:, :, None, None] + F.interpolate(low, (30, 54), mode="bilinear", align_corners=False)
    assert (_nchw(out_r.cpu()) - ref_r).abs().max() < 1e-5
    two = ops.cab(_nhwc(x).to(DEV), pooled, dp.to(DEV), ops.resize_bilinear(_nhwc(low).to(DEV), (30, 54)), w1.to(DEV), b1.to(DEV),
                  w2.to(DEV), b2.to(DEV))
    assert (two - out_r).abs().max().item() < 2e-6                      # resize + CAB as two kernels (fma contraction may differ)
    ref_v = x * gate[:, :, None, None] + dp[:, :, None, None]
    out_v = ops.cab(_nhwc(x).to(DEV), pooled, dp.to(DEV), dp.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    assert (_nchw(out_v.cpu()) - ref_v).abs().max() < 1e-5


@pytest.mark.parametrize("hw,hwd", [((30, 54), (15, 27)), ((120, 214), (60, 107)), ((15, 27), None), ((7, 9), (4, 5))])
def test_cab_gate_from_maps_is_bit_identical_to_the_separate_kernels(hw, hwd):
    """frtm_cab_gate_from_maps (both pools + the gate in two launches) == global_avgpool x2 + cab_gate, bit for bit; with
    ``hwd`` None the deeper pool is a given vector (the coarsest level of the refinement network)."""
    ops = _ops()
    from frtm_vos_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(61)
    B, C = 5, 64
    x = torch.randn(B, hw[0], hw[1], C, generator=g).to(DEV)
    w1, b1 = (torch.randn(64, 128, generator=g) / 11).to(DEV), torch.randn(64, generator=g).to(DEV)
    w2, b2 = (torch.randn(64, 64, generator=g) / 8).to(DEV), torch.randn(64, generator=g).to(DEV)
    sp = ops.global_avgpool(x)
    if hwd is None:
        deeper = torch.randn(B, C, generator=g).to(DEV)
        dp = deeper
    else:
        deeper = torch.randn(B, hwd[0], hwd[1], C, generator=g).to(DEV)
        dp = ops.global_avgpool(deeper)
    gate_ref = torch.empty((B, C), device=DEV)
    lib().cab_gate(ptr(sp), ptr(dp), ptr(w1), ptr(b1), ptr(w2), ptr(b2), B, C, ptr(gate_ref), stream())
    HWd = 0 if hwd is None else hwd[0] * hwd[1]
    nbytes = lib().cab_gate_from_maps_workspace(B, hw[0] * hw[1], HWd, C)
    ws = torch.empty(nbytes // 4, device=DEV)
    gate, pools = torch.empty((B, C), device=DEV), torch.empty((B, 2 * C), device=DEV)
    lib().cab_gate_from_maps(ptr(x), hw[0] * hw[1], C, None if hwd is None else ptr(deeper), HWd, C, ptr(dp) if hwd is None else None,
                             B, C, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(gate), ptr(pools), ptr(ws), nbytes, stream())
    assert torch.equal(pools[:, :C], sp) and torch.equal(pools[:, C:], dp)
    assert torch.equal(gate, gate_ref)
    # and through ops.cab with the pools left to it
    a = ops.cab(x, None, None if hwd is not None else dp, deeper, w1, b1, w2, b2)
    b = ops.cab(x, sp, dp, deeper, w1, b1, w2, b2)
    assert torch.equal(a, b)


def test_concat_helpers_and_layout():
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    src = torch.randn(2, 5, 7, 64, generator=g)
    out = torch.zeros(6, 5, 7, 68, device=DEV)
    ops.broadcast_objects(src.to(DEV), 3, out, channels=64)
    assert torch.equal(out[..., :64].cpu(), src.repeat_interleave(3, dim=0))
    sc = torch.randn(6, 5, 7, generator=g)
    ops.scatter_channel(sc.to(DEV), out, 64, 0)
    assert torch.equal(out[..., 64].cpu(), sc) and torch.all(out[..., 65:] == 0)
    y = ops.nhwc_to_nchw(out, channels=65)
    assert torch.equal(y.cpu(), out[..., :65].permute(0, 3, 1, 2).cpu())


def test_conv3x3_to1():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 32, 21, 33, generator=g)
    w = torch.randn(1, 32, 3, 3, generator=g) / 17
    b = torch.randn(1, generator=g)
    ref = F.conv2d(x, w, b, 1, 1)[:, 0]
    y = ops.conv3x3_to1(_nhwc(x).to(DEV), w.permute(2, 3, 1, 0).reshape(9, 32).contiguous().to(DEV), b.to(DEV))
    assert (y.cpu() - ref).abs().max() < 1e-5


def test_corr3x3_multi_filter():
    ops = _ops()
    g = torch.Generator().manual_seed(10)
    x = torch.randn(3, 96, 30, 54, generator=g)
    f = torch.randn(3, 96, 3, 3, generator=g) / 30
    idx = torch.tensor([2, 0, 1], dtype=torch.int32)
    y = ops.corr3x3(x.to(DEV), f.to(DEV), idx.to(DEV)).cpu()
    for n in range(3):
        ref = F.conv2d(x[n:n + 1], f[idx[n]:idx[n] + 1], None, 1, 1)[0, 0]
        assert (y[n] - ref).abs().max() < 2e-5


def test_pixel_weights_golden(golden):
    ops = _ops()
    import golden_inputs as GI
    g = golden("pw_merge")
    gen = torch.Generator().manual_seed(21)
    y = GI.blob_masks(6, GI.SMALL, gen, soft=False)
    y[4] = 0
    y[4, 0, 0, :5] = 1
    y[5] = (GI.blob_masks(1, GI.SMALL, gen)[0] * 0 + 1)
    y[5, 0, :8] = 0
    w = ops.pixel_weights(y.to(DEV), 0.1, False).cpu().numpy()
    assert np.abs(w - g["pw"]).max() < 1e-6
    # thresholded variant == weights of the binarised map
    soft = torch.rand(2, 1, *GI.SMALL, generator=gen)
    from oracle import frtm_ref as R
    w2 = ops.pixel_weights(soft.to(DEV), 0.1, True).cpu()
    assert (w2 - R.pixel_weights((soft > 0.5).float(), 0.1)).abs().max() < 1e-6


def test_merge_golden(golden):
    ops = _ops()
    import golden_inputs as GI
    g = golden("pw_merge")
    gen = torch.Generator().manual_seed(21)
    GI.blob_masks(6, GI.SMALL, gen, soft=False)
    GI.blob_masks(1, GI.SMALL, gen)
    probs = torch.rand(4, *GI.SMALL, generator=gen)
    probs[0] = 0
    probs[2, :10] = 1.0
    probs[3, -10:] = 0.0
    lut = torch.tensor([0, 3, 5, 9], dtype=torch.uint8)
    # logit_mask = 0: inputs are probabilities (start-mask path) -> exactly the golden case
    masks, labels, counts = ops.merge_masks(probs[1:].contiguous().to(DEV), 0, None, lut.to(DEV), False)
    assert np.abs(masks.cpu().numpy() - g["merged"]).max() < 1e-6
    assert np.array_equal(labels.cpu().numpy(), g["labels"])
    assert counts.cpu().tolist() == [int((g["merged"][i] > 0.5).sum()) for i in (1, 2, 3)]
    # logits path + suppression + single-object rule against the oracle restatement
    from oracle import frtm_ref as R
    lg = torch.randn(3, *GI.SMALL, generator=gen) * 4
    sup = (torch.rand(*GI.SMALL, generator=gen) > 0.9).to(torch.uint8)
    cm = torch.zeros(4, *GI.SMALL)
    cm[1:] = torch.sigmoid(lg) * (1 - sup).float()
    ref = R.merge_masks(cm)
    masks, labels, _ = ops.merge_masks(lg.to(DEV), 0b111, sup.to(DEV), lut.to(DEV), False)
    assert (masks.cpu() - ref).abs().max() < 2e-6
    ref_l = R.labels_from_masks(ref.clone(), lut, False)
    assert (labels.cpu() != ref_l).float().mean() < 1e-4       # ties at ulp level only
    cm1 = torch.zeros(2, *GI.SMALL)
    cm1[1] = torch.sigmoid(lg[0])
    ref1 = R.merge_masks(cm1)
    m1, l1, _ = ops.merge_masks(lg[:1].contiguous().to(DEV), 1, None, lut.to(DEV), True)
    assert (m1.cpu() - ref1).abs().max() < 2e-6
    assert (l1.cpu() != R.labels_from_masks(ref1.clone(), lut, True)[0]).float().mean() < 1e-4


@pytest.mark.parametrize("N", [1, 2, 5, 8, 9, 12])
def test_merge_object_counts(N):
    """Both dispatch branches of frtm_merge_masks (register kernel N <= 8, general kernel above) and the all-frames entry
    against the oracle restatement of tracker.py:208-221 / :143-150; ragged size (HW not a multiple of the block)."""
    ops = _ops()
    from oracle import frtm_ref as R
    gen = torch.Generator().manual_seed(100 + N)
    H, W, Fr = 37, 53, 3
    lg = torch.randn(Fr, N, H, W, generator=gen) * 1.5
    lg[:, :, :3] = 30.0                                     # saturated rows: every object at the clamp
    lg[:, :, -3:] = -30.0
    lut = torch.arange(0, 2 * (N + 1), 2, dtype=torch.uint8)
    single = N == 1
    counts = torch.full((Fr, N + 2), 7, dtype=torch.int32, device=DEV)
    masks_f, labels_f = ops.merge_masks_frames(lg.to(DEV), (1 << N) - 1, lut.to(DEV), single, counts)
    for f in range(Fr):
        cm = torch.zeros(N + 1, H, W)
        cm[1:] = torch.sigmoid(lg[f])
        ref = R.merge_masks(cm)
        ref_l = R.labels_from_masks(ref.clone(), lut, single)
        ref_l = ref_l[0] if ref_l.dim() == 3 else ref_l
        masks, labels, cnt = ops.merge_masks(lg[f].contiguous().to(DEV), (1 << N) - 1, None, lut.to(DEV), single)
        # the two entry points run the same arithmetic
        assert torch.equal(masks, masks_f[f]) and torch.equal(labels, labels_f[f])
        assert torch.equal(cnt, counts[f, :N])
        # away from saturation (p -> 1 makes p/(1-p) ill-conditioned, cf. test_gpu_model) the merged scores agree to rounding
        cond = (cm[1:].max(0).values < 0.9)
        assert ((masks.cpu() - ref).abs() * cond).max() < 2e-5
        assert ((labels.cpu() != ref_l) & cond).float().mean() < 1e-3
        # the count is the exact integer count of the masks the kernel wrote
        assert cnt.cpu().tolist() == [int((masks[i + 1] > 0.5).sum()) for i in range(N)]
        # invariants: one-hot support, labels consistent with the written masks
        assert int(((masks > 0).sum(0) > 1).sum()) == 0


def test_final_conv_commutes_with_upsampling():
    """conv3x3_to1_upsampled == conv3x3_to1(resize(pyrup(x))) (the reference order, seg_network.py:141-145)."""
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 32, 24, 43, generator=g)
    w = torch.randn(1, 32, 3, 3, generator=g) / 17
    b = torch.randn(1, generator=g)
    from oracle import frtm_ref as R
    size = (48, 84)
    ref = F.conv2d(F.interpolate(R.pyr_up_bicubic(x), size, mode="bilinear", align_corners=False), w, b, 1, 1)[:, 0]
    w9c = w.permute(2, 3, 1, 0).reshape(9, 32).contiguous().to(DEV)
    y = ops.conv3x3_to1_upsampled(_nhwc(x).to(DEV), w9c, b.to(DEV), size)
    assert (y.cpu() - ref).abs().max() < 1e-5


@pytest.mark.parametrize("hw,size", [((24, 43), (48, 84)), ((24, 43), (48, 86)), ((37, 50), (73, 99)), ((120, 214), (240, 427))])
def test_fused_upsampler_tail(hw, size):
    """frtm_upsample_tapsum (bicubic x2 -> bilinear -> 9 shifted taps in one kernel) == conv3x3_to1(resize(pyrup(x)))."""
    ops = _ops()
    from frtm_vos_b200._lib import lib
    g = torch.Generator().manual_seed(21)
    h, w = hw
    x = torch.randn(2, 32, h, w, generator=g)
    wt = torch.randn(1, 32, 3, 3, generator=g) / 17
    b = torch.randn(1, generator=g)
    from oracle import frtm_ref as R
    ref = F.conv2d(F.interpolate(R.pyr_up_bicubic(x), size, mode="bilinear", align_corners=False), wt, b, 1, 1)[:, 0]
    w9c = wt.permute(2, 3, 1, 0).reshape(9, 32).contiguous().to(DEV)
    xd = _nhwc(x).to(DEV)
    t12 = torch.empty((2, h, w, 12), device=DEV)
    lib().tapmaps_nhwc(xd.data_ptr(), 2 * h * w, 32, w9c.data_ptr(), t12.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert lib().upsample_tapsum_supported(h, w, size[0], size[1]) == 1
    y = ops.upsample_tapsum(t12, b.to(DEV), size)
    assert (y.cpu() - ref).abs().max() < 2e-5
    # and the unfused chain agrees too
    y2 = ops.conv3x3_to1_upsampled(xd, w9c, b.to(DEV), size)
    assert (y.cpu() - y2.cpu()).abs().max() < 2e-5


def test_pyrup_bicubic_split_planes():
    ops = _ops()
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, 9, 11, 64, generator=g).to(DEV)
    y = ops.pyrup_bicubic(x)
    sp = ops.pyrup_bicubic(x, split=True)
    rec = (sp.hi.float() + sp.lo.float()) / 16.0
    assert (rec - y).abs().max() < 2e-6 * max(y.abs().max().item(), 1.0)


@pytest.mark.parametrize("hw", [(64, 112), (33, 47), (480, 854)])
def test_stem_as_tensor_core_conv_over_patches(hw):
    """7x7/s2/p3 stem + folded BN + ReLU: im2col patches (frtm_stem_patches_u8) + 1x1 tcgen05 conv vs torch on the
    normalised image (feature_extractor.py:27-32,42; torchvision resnet.py conv1/bn1/relu)."""
    ops = _ops()
    g = torch.Generator().manual_seed(hw[0])
    B = 2
    img = torch.randint(0, 256, (B, 3, *hw), generator=g, dtype=torch.uint8)
    w = torch.randn(64, 3, 7, 7, generator=g) / 12
    bn = dict(weight=torch.rand(64, generator=g) + 0.5, bias=torch.randn(64, generator=g) * 0.1,
              running_mean=torch.randn(64, generator=g) * 0.1, running_var=torch.rand(64, generator=g) + 0.5)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = img.float() * (1 / 255 / std) + (-mean / std)
    y = F.conv2d(x.double(), w.double(), None, 2, 3)
    sc = bn["weight"].double() / torch.sqrt(bn["running_var"].double() + 1e-5)
    ref = torch.relu(y * sc.view(1, -1, 1, 1) + (bn["bias"].double() - bn["running_mean"].double() * sc).view(1, -1, 1, 1)).float()
    pc = ops.pack_conv_tc(ops.stem_weight_as_1x1(w), bn=bn, device=DEV)
    out = ops.conv2d_tc(ops.stem_patches(img.to(DEV)), pc, relu=True)["y"]
    assert out.shape == (B, (hw[0] - 1) // 2 + 1, (hw[1] - 1) // 2 + 1, 64)
    assert (out.cpu().permute(0, 3, 1, 2) - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max().item())


def test_host_wait_policy_applies_to_live_context():
    """parallel.set_host_wait_policy changes the primary context's scheduling flags in place (what bench.py relies on)."""
    from frtm_vos_b200.parallel import set_host_wait_policy
    torch.zeros(1, device=DEV)                      # context exists
    try:
        assert set_host_wait_policy(0, "yield") & 0x7 == 2
        assert set_host_wait_policy(0, "spin") & 0x7 == 1
    finally:
        assert set_host_wait_policy(0, "auto") & 0x7 == 0
    x = torch.arange(10, device=DEV).sum().item()   # synchronising call still works
    assert x == 45
