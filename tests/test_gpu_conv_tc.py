"""-m gpu: the tcgen05 tensor-core convolution (split-fp16, 3 MMAs per k-step) against torch fp32 on the CPU and
against the library's own exact fp32 CUDA-core conv."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("cin,cout,k,hw,B", [
    (64, 64, 1, (8, 16), 1),        # exactly one tile, one k-block
    (64, 64, 3, (8, 16), 1),        # 9 taps, halo entirely out of bounds
    (64, 64, 3, (30, 54), 2),       # ragged tiles, batch
    (128, 128, 3, (15, 27), 1),     # two k-chunks, BN=128
    (256, 256, 3, (30, 54), 1),     # two N tiles
    (64, 32, 3, (40, 44), 1),       # BN=32
    (64, 65, 3, (15, 27), 3),       # BN=80, ragged Cout
    (1024, 256, 1, (30, 54), 1),    # deep 1x1
    (256, 1024, 1, (9, 5), 1),      # W < tile width
])
def test_conv2d_tc_matches_fp32(cin, cout, k, hw, B):
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(cin + 7 * cout + k)
    x = torch.randn(B, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, 1, k // 2)
    res = torch.randn(ref.shape, generator=g)
    ref2 = F.relu(ref + res)
    xs = ops.split_f16(_nhwc(x).to(DEV))
    pc = ops.pack_conv_tc(w, b, device=DEV)
    o = ops.conv2d_tc(xs, pc, nchw=True, out_split=True)
    scale = max(1.0, ref.abs().max().item())
    err = (_nchw(o["y"].cpu()) - ref).abs().max().item()
    assert err < 1e-5 * scale, err
    assert (o["nchw"].cpu() - ref).abs().max().item() < 1e-5 * scale
    sp = o["split"]
    back = (sp.hi.float() + sp.lo.float())[..., :cout].cpu() / ops.ACT_SCALE
    assert (_nchw(back) - ref).abs().max().item() < 1e-5 * scale
    o2 = ops.conv2d_tc(xs, pc, res=_nhwc(res).to(DEV), relu=True)
    assert (_nchw(o2["y"].cpu()) - ref2).abs().max().item() < 1e-5 * scale
    if cout % 4 == 0:
        o3 = ops.conv2d_tc(xs, pc, res=ops.split_f16(_nhwc(res).to(DEV)), relu=True)
        assert (_nchw(o3["y"].cpu()) - ref2).abs().max().item() < 2e-5 * scale
    # and against the exact CUDA-core kernel of the same library
    y_simt = ops.conv2d(_nhwc(x).to(DEV), ops.pack_conv(w, b, device=DEV))
    assert (y_simt - o["y"]).abs().max().item() < 1e-5 * scale


def test_conv2d_tc_small_magnitudes_and_wide_input():
    """Weights of very different magnitude per output channel and tiny activations (subnormal-lo regime)."""
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 64, 16, 32, generator=g) * 1e-3
    w = torch.randn(64, 64, 3, 3, generator=g) * torch.logspace(-6, 1, 64).view(-1, 1, 1, 1)
    ref = F.conv2d(x.double(), w.double(), None, 1, 1)
    wide = torch.zeros(1, 16, 32, 72)
    wide[..., :64] = _nhwc(x)
    xs = ops.split_f16(wide.to(DEV), channels=64, ld=72)
    o = ops.conv2d_tc(xs, ops.pack_conv_tc(w, None, device=DEV))
    rel = ((_nchw(o["y"].cpu()).double() - ref).abs().amax(dim=(0, 2, 3)) / ref.abs().amax(dim=(0, 2, 3))).max().item()
    assert rel < 2e-5, rel


@pytest.mark.parametrize("cin,cout,k,hw", [(64, 128, 3, (120, 214)), (128, 256, 3, (61, 107)), (64, 128, 1, (60, 107)),
                                           (256, 512, 1, (30, 54)), (128, 128, 3, (16, 32))])
def test_conv2d_tc_stride2(cin, cout, k, hw):
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    ref = F.conv2d(x, w, None, 2, k // 2)
    o = ops.conv2d_tc(ops.split_f16(_nhwc(x).to(DEV)), ops.pack_conv_tc(w, None, device=DEV, stride=2), relu=False)
    assert o["y"].shape[1:3] == ref.shape[2:]
    assert (_nchw(o["y"].cpu()) - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cout,n_obj", [(65, 1), (65, 3), (64, 2)])
def test_conv65(cout, n_obj):
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(cout + n_obj)
    Fr, hw = 2, (15, 27)
    B = Fr * n_obj
    h = torch.randn(Fr, 64, *hw, generator=g)
    s = torch.randn(B, 1, *hw, generator=g)
    w = torch.randn(cout, 65, 3, 3, generator=g) / 24
    b = torch.randn(cout, generator=g)
    x = torch.cat((h.repeat_interleave(n_obj, 0), s), 1)
    ref = F.relu(F.conv2d(x, w, b, 1, 1))
    y, sp, extra = ops.conv65(ops.split_f16(_nhwc(h).to(DEV)), s[:, 0].contiguous().to(DEV), ops.pack_conv65(w, b, device=DEV),
                              n_obj=n_obj, want_f32=True)
    assert (_nchw(y.cpu()) - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    back = (sp.hi.float() + sp.lo.float()).cpu() / ops.ACT_SCALE
    assert (_nchw(back) - ref[:, :64]).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    if cout == 65:
        assert (extra.cpu() - ref[:, 64]).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cout,hw", [(65, (15, 27)), (64, (15, 27)), (65, (21, 40))])
def test_conv65_fused_epilogue(cout, hw):
    """n_obj == 1: the score channel's rank-1 term, bias and ReLU run inside the tensor-core conv's epilogue."""
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(7 * cout + hw[0])
    B = 3
    h = torch.randn(B, 64, *hw, generator=g)
    s = torch.randn(B, 1, *hw, generator=g)
    w = torch.randn(cout, 65, 3, 3, generator=g) / 24
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(torch.cat((h, s), 1), w, b, 1, 1))
    y, sp, extra = ops.conv65(ops.split_f16(_nhwc(h).to(DEV)), s[:, 0].contiguous().to(DEV), ops.pack_conv65(w, b, device=DEV))
    assert y is None and sp.hi.shape[-1] == 64
    back = (sp.hi.float() + sp.lo.float()).cpu() / ops.ACT_SCALE
    assert (_nchw(back) - ref[:, :64]).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    if cout == 65:
        assert (extra.cpu() - ref[:, 64]).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    else:
        assert extra is None


def test_device_packer_matches_host_packer():
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(4)
    for cout, cin in [(96, 256), (288, 1024), (65, 64)]:
        w = torch.randn(cout, cin, 1, 1, generator=g) * torch.logspace(-3, 1, cout).view(-1, 1, 1, 1)
        w[3] = 0
        a = ops.pack_conv_tc(w, None, device=DEV)
        b = ops.pack_conv_tc_1x1_device(w.to(DEV))
        assert a.bn == b.bn and torch.equal(a.oscale, b.oscale)
        assert torch.equal(a.wt.view(torch.int16), b.wt.view(torch.int16))


def test_conv2d_tc_tap_contraction_in_epilogue():
    """'tap' output: the 9 tap maps of a following 3x3 -> 1 conv, contracted per pixel from the ReLU'd conv output."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200._lib import lib
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 64, 20, 37, generator=g)
    w = torch.randn(32, 64, 3, 3, generator=g) / 24
    b = torch.randn(32, generator=g) * 0.1
    w2 = torch.randn(1, 32, 3, 3, generator=g) / 17
    pc = ops.pack_conv_tc(w, b, device=DEV)
    xs = ops.split_f16(x.permute(0, 2, 3, 1).contiguous().to(DEV))
    w9c = w2.permute(2, 3, 1, 0).reshape(9, 32).contiguous().to(DEV)
    o = ops.conv2d_tc(xs, pc, relu=True, tapw=w9c)
    y = torch.relu(F.conv2d(x.double(), w.double(), b.double(), 1, 1)).float()
    assert (o["y"].cpu().permute(0, 3, 1, 2) - y).abs().max() < 1e-5
    ref = torch.einsum("bchw,tc->bhwt", y.double(), w9c.cpu().double()).float()
    assert (o["tap"][..., :9].cpu() - ref).abs().max() < 1e-5
    assert float(o["tap"][..., 9:].abs().max()) == 0.0
    only = ops.conv2d_tc(xs, pc, relu=True, out_f32=False, tapw=w9c)
    assert only["y"] is None and torch.equal(only["tap"], o["tap"])


@pytest.mark.parametrize("cout,hw,B", [(64, (8, 16), 1), (64, (30, 54), 2), (64, (37, 29), 3), (32, (40, 44), 1), (64, (120, 214), 24)])
def test_slab_kernel_matches_general_kernel(cout, hw, B):
    """3x3 / stride 1 / 64 input channels: the persistent slab kernel (resident weights, column-shifted slabs) against
    the general tensor-core kernel and against fp32 conv2d; more tiles than SMs in the last case."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200._lib import lib
    g = torch.Generator().manual_seed(cout + hw[0])
    x = torch.randn(B, 64, *hw, generator=g)
    w = torch.randn(cout, 64, 3, 3, generator=g) / 24
    b = torch.randn(cout, generator=g)
    res = torch.randn(B, cout, *hw, generator=g)
    xs = ops.split_f16(_nhwc(x).to(DEV))
    pc = ops.pack_conv_tc(w, b, device=DEV)
    rs = _nhwc(res).to(DEV)
    out = {}
    for on in (1, 0):
        # kernel_select is a per-call argument of frtm_conv2d_tc: 0 = library's choice (the specialised kernel), 1 = general
        out[on] = ops.conv2d_tc(xs, pc, res=rs, relu=True, out_split=True, nchw=(B <= 3), kernel_select=1 - on)
        torch.cuda.synchronize()
    if B <= 3:
        ref = F.relu(F.conv2d(x, w, b, 1, 1) + res)
        scale = max(1.0, ref.abs().max().item())
        assert (_nchw(out[1]["y"].cpu()) - ref).abs().max().item() < 1e-5 * scale
        assert (out[1]["nchw"].cpu() - ref).abs().max().item() < 1e-5 * scale
    scale = max(1.0, out[0]["y"].abs().max().item())
    assert (out[1]["y"] - out[0]["y"]).abs().max().item() < 4e-6 * scale
    d = (out[1]["split"].hi.float() + out[1]["split"].lo.float()) - (out[0]["split"].hi.float() + out[0]["split"].lo.float())
    assert d.abs().max().item() / 16.0 < 4e-6 * scale


@pytest.mark.parametrize("cin,cout,hw,B", [(64, 64, (8, 16), 1), (64, 64, (37, 29), 3), (128, 64, (30, 54), 2), (192, 64, (33, 47), 2),
                                           (256, 32, (15, 27), 2), (64, 64, (120, 214), 24)])
def test_streaming_1x1_kernel_matches_general_kernel(cin, cout, hw, B):
    """1x1 / stride 1 / Cin <= 256 / one N tile: the persistent streaming kernel against the general kernel and fp32."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200._lib import lib
    g = torch.Generator().manual_seed(cin + cout + hw[0])
    x = torch.randn(B, cin, *hw, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(B, cout, *hw, generator=g)
    xs = ops.split_f16(_nhwc(x).to(DEV))
    pc = ops.pack_conv_tc(w, b, device=DEV)
    rs = _nhwc(res).to(DEV)
    out = {}
    for on in (1, 0):
        # kernel_select is a per-call argument of frtm_conv2d_tc: 0 = library's choice (the specialised kernel), 1 = general
        out[on] = ops.conv2d_tc(xs, pc, res=rs, relu=True, out_split=True, nchw=(B <= 3), kernel_select=1 - on)
        torch.cuda.synchronize()
    if B <= 3:
        ref = F.relu(F.conv2d(x, w, b) + res)
        scale = max(1.0, ref.abs().max().item())
        assert (_nchw(out[1]["y"].cpu()) - ref).abs().max().item() < 1e-5 * scale
        assert (out[1]["nchw"].cpu() - ref).abs().max().item() < 1e-5 * scale
    scale = max(1.0, out[0]["y"].abs().max().item())
    assert (out[1]["y"] - out[0]["y"]).abs().max().item() < 4e-6 * scale
    d = (out[1]["split"].hi.float() + out[1]["split"].lo.float()) - (out[0]["split"].hi.float() + out[0]["split"].lo.float())
    assert d.abs().max().item() / 16.0 < 4e-6 * scale


def test_conv2d_tc_wide_tile_128():
    """The N = 128 tile (three stages, all 512 TMEM columns) is no longer the default for wide convs; keep it covered."""
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(128)
    x = torch.randn(2, 128, 15, 27, generator=g)
    w = torch.randn(256, 128, 3, 3, generator=g) / (128 * 9) ** 0.5
    b = torch.randn(256, generator=g)
    ref = F.relu(F.conv2d(x, w, b, 1, 1))
    xs = ops.split_f16(_nhwc(x).to(DEV))
    for tile in (128, 64):
        pc = ops.pack_conv_tc(w, b, device=DEV, bn_tile=tile)
        assert pc.bn == tile
        o = ops.conv2d_tc(xs, pc, relu=True)
        assert (_nchw(o["y"].cpu()) - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,stride,hw,B", [
    (128, 128, 3, 1, (15, 27), 1),      # pair N = 128 only, 2 tiles = one pair
    (256, 256, 3, 1, (30, 54), 1),      # rn101 stage-3 3x3 (one frame)
    (256, 256, 3, 1, (30, 54), 8),      # ... at the block batch: 128 tiles, N = 256 chosen automatically
    (1024, 256, 1, 1, (30, 54), 8),     # rn101 stage-3 reduce
    (256, 1024, 1, 1, (30, 54), 8),     # rn101 stage-3 expand (residual + ReLU epilogue)
    (512, 1024, 1, 2, (60, 107), 2),    # stage-3 downsample, stride 2
    (256, 256, 3, 2, (60, 107), 3),     # stride-2 3x3, odd tile count (last pair half empty)
    (512, 512, 3, 1, (15, 27), 8),      # rn18 layer4 / rn101 stage 4
    (128, 512, 1, 1, (60, 107), 1),     # stage-2 expand
])
def test_pair_kernel_matches_fp32_and_general_kernel(cin, cout, k, stride, hw, B):
    """conv_tc2_kernel (tcgen05.mma.cta_group::2 over a CTA pair, M = 256 x N = 128 / 256) at the ResNet-101 / ResNet-18
    production shapes: against torch fp32 on the CPU and against the general tile kernel, for both pair tile widths."""
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(cin + 3 * cout + k + stride)
    x = torch.randn(B, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, stride, k // 2)
    res = torch.randn(ref.shape, generator=g)
    ref = F.relu(ref + res)
    scale = max(1.0, ref.abs().max().item())
    xs = ops.split_f16(_nhwc(x).to(DEV))
    pc = ops.pack_conv_tc(w, b, device=DEV, stride=stride)
    assert pc.bn == 64
    rs = ops.split_f16(_nhwc(res).to(DEV))
    gen = ops.conv2d_tc(xs, pc, res=rs, relu=True, out_split=True, kernel_select=1)
    sels = [0, 2, 4] + ([3] if cout % 256 == 0 else [])
    for sel in sels:
        o = ops.conv2d_tc(xs, pc, res=rs, relu=True, out_split=True, nchw=True, kernel_select=sel)
        torch.cuda.synchronize()
        assert (_nchw(o["y"].cpu()) - ref).abs().max().item() < 1e-5 * scale, sel
        assert (o["nchw"].cpu() - ref).abs().max().item() < 1e-5 * scale, sel
        assert (o["y"] - gen["y"]).abs().max().item() < 4e-6 * scale, sel
        d = (o["split"].hi.float() + o["split"].lo.float()) - (gen["split"].hi.float() + gen["split"].lo.float())
        assert d.abs().max().item() / 16.0 < 4e-6 * scale, sel


@pytest.mark.parametrize("hw,B", [((480, 854), 2), ((128, 224), 3), ((61, 75), 1), ((720, 1280), 1)])
def test_fused_stem_is_bit_identical_to_patches_plus_conv(hw, B):
    """frtm_stem_conv_u8 (im2col patches of the 7x7/s2 stem built in shared memory) against frtm_stem_patches_u8 +
    frtm_conv2d_tc bit for bit, and against F.conv2d on the normalised image (torchvision resnet.py conv1 + bn1 + relu)."""
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(hw[0] + B)
    img = torch.randint(0, 256, (B, 3, *hw), dtype=torch.uint8, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) / 12
    bn = dict(weight=torch.rand(64, generator=g) + 0.5, bias=torch.randn(64, generator=g) * 0.1,
              running_mean=torch.randn(64, generator=g) * 0.1, running_var=torch.rand(64, generator=g) + 0.5)
    pc = ops.pack_conv_tc(ops.stem_weight_as_1x1(w), bn=bn, device=DEV)
    d = img.to(DEV)
    old = ops.conv2d_tc(ops.stem_patches(d), pc, relu=True)["y"]
    new = ops.stem_conv(d, pc)
    torch.cuda.synchronize()
    assert torch.equal(old, new)
    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = (img.float() / 255 - mean) / std
    ref = F.conv2d(x, w, None, 2, 3)
    ref = (ref - bn["running_mean"].view(1, -1, 1, 1)) / torch.sqrt(bn["running_var"].view(1, -1, 1, 1) + 1e-5) * \
        bn["weight"].view(1, -1, 1, 1) + bn["bias"].view(1, -1, 1, 1)
    ref = F.relu(ref)
    assert (_nchw(new.cpu()) - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())

