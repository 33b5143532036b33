"""Thin tensor-level wrappers over the C ABI (``include/frtm_b200.h``).

PyTorch is used for device memory and streams only; every function here enqueues hand-written kernels from
``libfrtm_b200.so`` on the current CUDA stream and returns the output tensor(s).  Activations are NHWC fp32
(``(B,H,W,ld)`` with the logical channel count carried separately when ``ld`` is padded).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from ._lib import lib, ptr, stream, require_cuda


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class PackedConv:
    """Conv weights in the library layout ``[kh][kw*cin_pad][cout_pad]`` (+ optional bias with BN folded in)."""
    w: torch.Tensor
    bias: Optional[torch.Tensor]
    cin: int
    cin_pad: int
    cout: int
    kh: int
    kw: int
    stride: int
    pad: int

    @property
    def flops(self) -> int:
        return 2 * self.cin * self.cout * self.kh * self.kw


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn: Optional[dict] = None, stride: int = 1,
              pad: Optional[int] = None, device=None, cin_pad: Optional[int] = None, eps: float = 1e-5) -> PackedConv:
    """(Cout,Cin,kh,kw) [+ eval-mode BatchNorm folded] -> PackedConv on ``device``.  Host-side, done once per model."""
    w = weight.detach().to("cpu", torch.float64)
    cout, cin, kh, kw = w.shape
    b = bias.detach().to("cpu", torch.float64) if bias is not None else None
    if bn is not None:
        scale = bn["weight"].double().cpu() / torch.sqrt(bn["running_var"].double().cpu() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        shift = bn["bias"].double().cpu() - bn["running_mean"].double().cpu() * scale
        b = shift if b is None else b * scale + shift
    cin_pad = cin_pad or _rup(cin, 4)
    cout_pad = _rup(cout, 4)
    packed = torch.zeros(kh, kw, cin_pad, cout_pad, dtype=torch.float64)
    packed[:, :, :cin, :cout] = w.permute(2, 3, 1, 0)
    packed = packed.reshape(kh, kw * cin_pad, cout_pad).to(torch.float32).contiguous().to(device)
    return PackedConv(packed, None if b is None else b.to(torch.float32).contiguous().to(device), cin, cin_pad, cout, kh,
                      kw, stride, kh // 2 if pad is None else pad)


def conv2d(x: torch.Tensor, pc: PackedConv, res: Optional[torch.Tensor] = None, relu: bool = False,
           out: Optional[torch.Tensor] = None, coff: int = 0, nchw: bool = False, nhwc: bool = True):
    """x (B,H,W,ldx) -> y (B,Ho,Wo,cout) [written into ``out`` at channel offset ``coff`` when given].

    Returns ``y`` or ``(y, y_nchw)`` when ``nchw``; with ``nhwc=False`` only the NCHW tensor is produced."""
    B, H, W, ldx = x.shape
    Ho = (H + 2 * pc.pad - pc.kh) // pc.stride + 1
    Wo = (W + 2 * pc.pad - pc.kw) // pc.stride + 1
    y = None
    if nhwc:
        y = out if out is not None else torch.empty((B, Ho, Wo, pc.cout), device=x.device, dtype=torch.float32)
        assert y.shape[:3] == (B, Ho, Wo) and y.is_contiguous()
    y_nchw = torch.empty((B, pc.cout, Ho, Wo), device=x.device, dtype=torch.float32) if nchw else None
    if res is not None:
        assert res.shape[:3] == (B, Ho, Wo)
    lib().conv2d_nhwc(ptr(x), B, H, W, pc.cin_pad, ldx, ptr(pc.w), ptr(pc.bias), ptr(res),
                      0 if res is None else res.shape[3], ptr(y), 0 if y is None else y.shape[3], coff, ptr(y_nchw),
                      pc.cout, pc.kh, pc.kw, pc.stride, pc.pad, 1 if relu else 0, stream())
    if nchw and nhwc:
        return y, y_nchw
    return y_nchw if nchw else y


def normalize_u8(img: torch.Tensor) -> torch.Tensor:
    """uint8 (B,3,H,W) -> fp32 NHWC (B,H,W,4)."""
    require_cuda(img, "image")
    B, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.uint8
    out = torch.empty((B, H, W, 4), device=img.device, dtype=torch.float32)
    lib().normalize_u8(ptr(img.contiguous()), B, H, W, ptr(out), stream())
    return out


def maxpool3x3s2(x: torch.Tensor, nchw: bool = False):
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
    y_nchw = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32) if nchw else None
    lib().maxpool3x3s2_nhwc(ptr(x), B, H, W, C, ptr(y), ptr(y_nchw), stream())
    return (y, y_nchw) if nchw else y


def maxpool3x3s2_split(x: torch.Tensor, want_f32: bool = False):
    """3x3/s2/p1 max pooling straight into split planes; returns (Split, fp32 y or None)."""
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32) if want_f32 else None
    sp = Split(torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float16),
               torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float16), C)
    lib().maxpool3x3s2_split_nhwc(ptr(x), B, H, W, C, ptr(y), ptr(sp.hi), ptr(sp.lo), stream())
    return sp, y


def resize_bilinear(x: torch.Tensor, size, channels: Optional[int] = None, out: Optional[torch.Tensor] = None,
                    coff: int = 0, accumulate: bool = False) -> torch.Tensor:
    B, H, W, ldx = x.shape
    C = channels or ldx
    Ho, Wo = int(size[0]), int(size[1])
    y = out if out is not None else torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
    lib().resize_bilinear_nhwc(ptr(x), B, H, W, C, ldx, ptr(y), Ho, Wo, y.shape[3], coff, 1 if accumulate else 0, stream())
    return y


def resize_bicubic(x: torch.Tensor, size, channels: Optional[int] = None) -> torch.Tensor:
    """F.interpolate(mode='bicubic', align_corners=False) of an NHWC tensor (C % 4 == 0)."""
    B, H, W, ldx = x.shape
    C = channels or ldx
    Ho, Wo = int(size[0]), int(size[1])
    y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
    lib().resize_bicubic_nhwc(ptr(x), B, H, W, C, ldx, ptr(y), Ho, Wo, C, stream())
    return y


def stem_weight_as_1x1(weight: torch.Tensor) -> torch.Tensor:
    """conv1.weight (64,3,7,7) -> (64,192,1,1) in the k order of ``stem_patches``: k = ky*24 + c*8 + kx."""
    cout = weight.shape[0]
    w = torch.zeros(cout, 7, 3, 8, dtype=weight.dtype)
    w[..., :7] = weight.detach().cpu().permute(0, 2, 1, 3)
    out = torch.zeros(cout, 192, dtype=weight.dtype)
    out[:, :168] = w.reshape(cout, 168)
    return out.reshape(cout, 192, 1, 1)


def stem_patches(images: torch.Tensor) -> "Split":
    """uint8 (B,3,H,W) -> Split planes (B,Ho,Wo,192): the 7x7/s2/p3 patches of the normalised image (k = ky*24+c*8+kx)."""
    B, C, H, W = images.shape
    assert C == 3 and images.dtype == torch.uint8
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    sp = Split(torch.empty((B, Ho, Wo, 192), device=images.device, dtype=torch.float16),
               torch.empty((B, Ho, Wo, 192), device=images.device, dtype=torch.float16), 192)
    lib().stem_patches_u8(ptr(images.contiguous()), B, H, W, ptr(sp.hi), ptr(sp.lo), stream())
    return sp


def stem_conv(images: torch.Tensor, pc: "PackedConvTC", relu: bool = True) -> torch.Tensor:
    """uint8 (B,3,H,W) -> fp32 NHWC (B,Ho,Wo,64): the 7x7/s2/p3 stem (+ folded BatchNorm + ReLU) in one kernel; the im2col
    patches are built in shared memory.  ``pc`` = ``pack_conv_tc(stem_weight_as_1x1(conv1.weight), bn=...)``.  Bit-identical
    to ``conv2d_tc(stem_patches(images), pc, relu=relu)``."""
    B, C, H, W = images.shape
    assert C == 3 and images.dtype == torch.uint8 and pc.cout == 64 and pc.cin == 192 and pc.bn == 64 and pc.k == 1
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, 64), device=images.device, dtype=torch.float32)
    lib().stem_conv_u8(ptr(images.contiguous()), B, H, W, ptr(pc.wt), ptr(pc.oscale), ptr(pc.bias), ptr(y), 64, 1 if relu else 0,
                       stream())
    return y


def pyrup_bicubic(x: torch.Tensor, split: bool = False):
    """x2 bicubic pyramid upsample; ``split=True`` returns the result as ``Split`` planes (the next conv's input format)
    without ever writing the fp32 tensor."""
    B, H, W, C = x.shape
    if split:
        assert C % 8 == 0
        sp = Split(torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.float16),
                   torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.float16), C)
        lib().pyrup_bicubic_nhwc(ptr(x), B, H, W, C, None, ptr(sp.hi), ptr(sp.lo), stream())
        return sp
    y = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.float32)
    lib().pyrup_bicubic_nhwc(ptr(x), B, H, W, C, ptr(y), None, None, stream())
    return y


def global_avgpool(x: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    B, H, W, ldx = x.shape
    C = channels or ldx
    out = torch.empty((B, C), device=x.device, dtype=torch.float32)
    nbytes = lib().global_avgpool_workspace(B, H * W, C)
    ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
    lib().global_avgpool_nhwc(ptr(x), B, H * W, C, ldx, ptr(out), ptr(ws), nbytes, stream())
    return out


def cab(shallow: torch.Tensor, shallow_pool: Optional[torch.Tensor], deep_pool: Optional[torch.Tensor], deeper: torch.Tensor,
        w1, b1, w2, b2, out_split: bool = False):
    """gate from pooled vectors, then shallow*gate + deeper (a map of the same shape, a lower-resolution map that is resized
    bilinearly on the fly, or a (B,C) vector).  ``shallow_pool`` / ``deep_pool`` = None: the pools are taken from the maps
    themselves (``deeper`` must be a map when ``deep_pool`` is None) — at C = 64 in two launches together with the gate."""
    B, H, W, C = shallow.shape
    gate = torch.empty((B, C), device=shallow.device, dtype=torch.float32)
    if shallow_pool is None and C == 64 and shallow.is_contiguous() and (deep_pool is not None or (deeper.dim() == 4 and deeper.is_contiguous())):
        from_map = deep_pool is None
        HWd = deeper.shape[1] * deeper.shape[2] if from_map else 0
        nbytes = lib().cab_gate_from_maps_workspace(B, H * W, HWd, C)
        ws = torch.empty(nbytes // 4, device=shallow.device, dtype=torch.float32)
        lib().cab_gate_from_maps(ptr(shallow), H * W, C, ptr(deeper) if from_map else None, HWd, C,
                                 None if from_map else ptr(deep_pool), B, C, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(gate), None,
                                 ptr(ws), nbytes, stream())
    else:
        if shallow_pool is None:
            shallow_pool = global_avgpool(shallow)
        if deep_pool is None:
            deep_pool = global_avgpool(deeper)
        lib().cab_gate(ptr(shallow_pool), ptr(deep_pool), ptr(w1), ptr(b1), ptr(w2), ptr(b2), B, C, ptr(gate), stream())
    if deeper.dim() == 4 and tuple(deeper.shape[1:3]) != (H, W):
        # deeper level at its own resolution: resized bilinearly inside the kernel (no full-size intermediate); with
        # ``out_split`` the result leaves as the split planes the next tensor-core conv reads (no fp32 round trip)
        assert deeper.shape[0] == B and deeper.shape[3] == C and deeper.is_contiguous()
        if out_split:
            sp = Split(torch.empty((B, H, W, C), device=shallow.device, dtype=torch.float16),
                       torch.empty((B, H, W, C), device=shallow.device, dtype=torch.float16), C)
            lib().cab_apply_resized_nhwc(ptr(shallow), ptr(gate), ptr(deeper), B, H, W, C, deeper.shape[1], deeper.shape[2], None,
                                         ptr(sp.hi), ptr(sp.lo), stream())
            return sp
        out = torch.empty_like(shallow)
        lib().cab_apply_resized_nhwc(ptr(shallow), ptr(gate), ptr(deeper), B, H, W, C, deeper.shape[1], deeper.shape[2], ptr(out),
                                     None, None, stream())
        return out
    out = torch.empty_like(shallow)
    lib().cab_apply_nhwc(ptr(shallow), ptr(gate), ptr(deeper), 1 if deeper.dim() == 2 else 0, B, H * W, C, ptr(out), stream())
    return split_f16(out) if out_split else out


def scatter_channel(src: torch.Tensor, dst: torch.Tensor, coff: int, nzero: int):
    """src (B,H,W) -> dst[..., coff], zeroing dst[..., coff+1 : coff+1+nzero]."""
    B, H, W, ld = dst.shape
    lib().scatter_channel_nhwc(ptr(src), B, H * W, ptr(dst), ld, coff, nzero, stream())


def broadcast_objects(src: torch.Tensor, n_obj: int, out: torch.Tensor, channels: Optional[int] = None):
    """src (F,H,W,lds) -> out (F*n_obj,H,W,ldd)[..., :C]."""
    F, H, W, lds = src.shape
    C = channels or lds
    lib().broadcast_objects_nhwc(ptr(src), F, n_obj, H * W, C, lds, ptr(out), out.shape[3], stream())


def nhwc_to_nchw(x: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    B, H, W, ld = x.shape
    C = channels or ld
    y = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
    lib().nhwc_to_nchw(ptr(x), B, H * W, C, ld, ptr(y), stream())
    return y


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Layout plumbing for foreign NCHW inputs (not on the fast path)."""
    return x.permute(0, 2, 3, 1).contiguous()


def conv3x3_to1(x: torch.Tensor, w9c: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    B, H, W, C = x.shape
    y = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    lib().conv3x3_to1_nhwc(ptr(x), B, H, W, C, ptr(w9c), ptr(bias), ptr(y), stream())
    return y


def merge_masks(src: torch.Tensor, logit_mask: int, suppress: Optional[torch.Tensor], lut: torch.Tensor, single: bool,
                masks: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None):
    """src (N,H,W) -> masks (N+1,H,W), labels (H,W) uint8, counts (N) int32."""
    N, H, W = src.shape
    masks = masks if masks is not None else torch.empty((N + 1, H, W), device=src.device, dtype=torch.float32)
    labels = torch.empty((H, W), device=src.device, dtype=torch.uint8)
    if counts is None:
        counts = torch.zeros(N, device=src.device, dtype=torch.int32)
    else:
        counts.zero_()
    lib().merge_masks(ptr(src), logit_mask, ptr(suppress), N, H * W, ptr(lut), 1 if single else 0, ptr(masks), ptr(labels),
                      ptr(counts), stream())
    return masks, labels, counts


def merge_masks_frames(src: torch.Tensor, logit_mask: int, lut: torch.Tensor, single: bool, counts: torch.Tensor):
    """src (F,N,H,W) -> masks (F,N+1,H,W), labels (F,H,W) uint8; counts (F,>=N) int32 is zeroed and filled."""
    F, N, H, W = src.shape
    masks = torch.empty((F, N + 1, H, W), device=src.device, dtype=torch.float32)
    labels = torch.empty((F, H, W), device=src.device, dtype=torch.uint8)
    counts.zero_()
    lib().merge_masks_frames(ptr(src), F, logit_mask, N, H * W, ptr(lut), 1 if single else 0, ptr(masks), ptr(labels),
                             ptr(counts), counts.shape[1], stream())
    return masks, labels


def sigmoid_suppress(logits: torch.Tensor, suppress: Optional[torch.Tensor]) -> torch.Tensor:
    """logits (N,H,W), suppress (H,W) uint8 or None -> sigmoid(logits) * (1 - suppress)."""
    N, H, W = logits.shape
    out = torch.empty_like(logits)
    lib().sigmoid_suppress(ptr(logits.contiguous()), ptr(suppress), N, H * W, ptr(out), stream())
    return out


def threshold(x: torch.Tensor, thr: float = 0.5) -> torch.Tensor:
    """(x > thr) as float32, same shape."""
    x = x.contiguous()
    out = torch.empty_like(x, dtype=torch.float32)
    lib().threshold_f32(ptr(x), x.numel(), float(thr), ptr(out), stream())
    return out


def labels_from_probs(probs: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """probs (F,N,H,W) raw object probabilities -> labels (F,H,W) uint8 (single-stage softmax/argmax rule)."""
    F, N, H, W = probs.shape
    labels = torch.empty((F, H, W), device=probs.device, dtype=torch.uint8)
    lib().labels_from_probs(ptr(probs.contiguous()), F, N, H * W, ptr(lut), ptr(labels), stream())
    return labels


def corr3x3(x: torch.Tensor, filt: torch.Tensor, index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (NB,c,h,w) NCHW, filt (NF,c,3,3) -> (NB,h,w); sample n uses filter index[n] (default 0)."""
    NB, c, h, w = x.shape
    out = torch.empty((NB, h, w), device=x.device, dtype=torch.float32)
    lib().corr3x3_nchw(ptr(x), ptr(filt), ptr(index), NB, c, h, w, ptr(out), stream())
    return out


def pixel_weights(y: torch.Tensor, tf: float, threshold: bool, return_count: bool = False, counts=None):
    """y (K,1,H,W) or (K,H,W) float -> hinge pixel weights of the same shape (and the per-map pixel counts).
    ``counts`` (int32 (K,), e.g. from merge_masks) skips the counting pass."""
    K = y.shape[0]
    HW = y.shape[-1] * y.shape[-2]
    w = torch.empty_like(y)
    ws = torch.empty(K, device=y.device, dtype=torch.float32) if counts is None else None
    lib().pixel_weights(ptr(y), K, HW, float(tf), 1 if threshold else 0, ptr(w), ptr(ws), ptr(counts), stream())
    return (w, ws) if return_count else w


def upsample_tapsum(t12: torch.Tensor, bias, image_size) -> torch.Tensor:
    """Tap maps (B,h,w,12) of the final 3x3 conv -> bicubic x2 -> bilinear to image_size -> shifted sum + bias -> (B,H,W),
    in one kernel when the sizes fit its shared-memory windows, else through the separate kernels."""
    B, h, w, _ = t12.shape
    H, W = int(image_size[0]), int(image_size[1])
    out = torch.empty((B, H, W), device=t12.device, dtype=torch.float32)
    if lib().upsample_tapsum_supported(h, w, H, W):
        lib().upsample_tapsum(ptr(t12), B, h, w, H, W, ptr(bias), ptr(out), stream())
    else:
        u = resize_bilinear(pyrup_bicubic(t12), (H, W))
        lib().shift_sum9(ptr(u), B, H, W, ptr(bias), ptr(out), stream())
    return out


def shift_sum9(t12: torch.Tensor, bias) -> torch.Tensor:
    """Full-resolution tap maps (B,H,W,12) -> sum of the 9 tap-shifted maps (zero padding) + bias -> (B,H,W)."""
    B, H, W, _ = t12.shape
    out = torch.empty((B, H, W), device=t12.device, dtype=torch.float32)
    lib().shift_sum9(ptr(t12), B, H, W, ptr(bias), ptr(out), stream())
    return out


def conv3x3_to1_upsampled(x: torch.Tensor, w9c: torch.Tensor, bias, image_size) -> torch.Tensor:
    """conv3x3_to1(resize_bilinear(pyrup_bicubic(x), image_size)) evaluated with the channel contraction first:
    9 tap maps at low resolution -> bicubic x2 -> bilinear -> shifted sum.  x (B,h,w,C) -> (B,H,W)."""
    B, h, w, C = x.shape
    t = torch.empty((B, h, w, 12), device=x.device, dtype=torch.float32)
    lib().tapmaps_nhwc(ptr(x), B * h * w, C, ptr(w9c), ptr(t), stream())
    u = resize_bilinear(pyrup_bicubic(t), image_size)
    H, W = int(image_size[0]), int(image_size[1])
    out = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    lib().shift_sum9(ptr(u), B, H, W, ptr(bias), ptr(out), stream())
    return out


def build_stencil(pw: torch.Tensor, y: torch.Tensor, fsize):
    """pw, y (K,1,H,W) -> stencil (K,9,h,w), uty (K,h,w)."""
    K = pw.shape[0]
    H, W = pw.shape[-2:]
    h, w = int(fsize[0]), int(fsize[1])
    st = torch.empty((K, 9, h, w), device=pw.device, dtype=torch.float32)
    uty = torch.empty((K, h, w), device=pw.device, dtype=torch.float32)
    lib().build_stencil(ptr(pw), ptr(y), K, H, W, h, w, ptr(st), ptr(uty), stream())
    return st, uty


# ---------------------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) convolution path: split-fp16 operands
# ---------------------------------------------------------------------------------------------------------------------
ACT_SCALE = 16.0


@dataclass
class Split:
    """Activation as two fp16 NHWC planes with hi + lo = ACT_SCALE * x (channel stride = hi.shape[3])."""
    hi: torch.Tensor
    lo: torch.Tensor
    channels: int

    @property
    def shape(self):
        return self.hi.shape


@dataclass
class PackedConvTC:
    wt: torch.Tensor          # fp16 [ntile][tap][cin/64][2][bn][64], 128-byte swizzled rows
    oscale: torch.Tensor      # fp32 [cout_pad]
    bias: Optional[torch.Tensor]
    cin: int
    cout: int
    k: int
    bn: int
    stride: int = 1


def _pick_bn(cout: int) -> int:
    """N tile of the tensor-core conv.  Wide convs use 64 too: measured on B200 (rn18 and rn101 at 480p) two resident CTAs
    per SM with N = 64 tiles beat one CTA per SM with N = 128 tiles (wave quantisation of the 56-tile stages, and the
    128-wide tile needs all 512 TMEM columns); FRTM_BN_MAX=128 restores the wide tile for experiments."""
    import os
    if cout <= 32:
        return 32
    if cout <= 64:
        return 64
    if cout <= 80:
        return 80
    return 128 if int(os.environ.get("FRTM_BN_MAX", "64")) >= 128 else 64


def pack_conv_tc(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn: Optional[dict] = None, device=None,
                 eps: float = 1e-5, bn_tile: Optional[int] = None, cin_pad: Optional[int] = None, stride: int = 1) -> PackedConvTC:
    """(Cout,Cin,k,k) [+ folded eval-mode BN] -> pre-split, pre-swizzled weight tiles for frtm_conv2d_tc (host, once)."""
    w = weight.detach().to("cpu", torch.float64)
    cout, cin, kh, kw = w.shape
    assert kh == kw and kh in (1, 3)
    b = bias.detach().to("cpu", torch.float64) if bias is not None else None
    if bn is not None:
        scale = bn["weight"].double().cpu() / torch.sqrt(bn["running_var"].double().cpu() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        shift = bn["bias"].double().cpu() - bn["running_mean"].double().cpu() * scale
        b = shift if b is None else b * scale + shift
    cin_p = cin_pad or _rup(cin, 64)
    assert cin_p % 64 == 0
    tile = bn_tile or _pick_bn(cout)
    ntile = (cout + tile - 1) // tile
    cout_p = ntile * tile
    # per-output-channel power-of-two scale: max |w| -> [2^9, 2^10)
    amax = w.abs().amax(dim=(1, 2, 3))
    expo = torch.where(amax > 0, 9 - torch.floor(torch.log2(amax.clamp_min(1e-300))), torch.zeros_like(amax))
    wscale = torch.pow(torch.tensor(2.0, dtype=torch.float64), expo)
    ws = torch.zeros(cout_p, cin_p, kh, kw, dtype=torch.float64)
    ws[:cout, :cin] = w * wscale.view(-1, 1, 1, 1)
    hi = ws.to(torch.float16)
    lo = (ws - hi.double()).to(torch.float16)
    oscale = torch.ones(cout_p, dtype=torch.float64)
    oscale[:cout] = 1.0 / (ACT_SCALE * wscale)
    # [2][cout_p][cin_p][kh][kw] -> [ntile][tap][kc][2][tile][64]
    t = torch.stack((hi, lo))                                          # (2, cout_p, cin_p, kh, kw)
    t = t.reshape(2, ntile, tile, cin_p // 64, 64, kh * kw)            # (2, nt, n, kc, k, tap)
    t = t.permute(1, 5, 3, 0, 2, 4).contiguous()                       # (nt, tap, kc, 2, n, k)
    # 128-byte swizzle: 16-byte chunk j of row r is stored at chunk position j ^ (r % 8)
    t = t.reshape(ntile, kh * kw, cin_p // 64, 2, tile // 8, 8, 8, 8)   # (..., r8, r, chunk, elem)
    sw = torch.empty_like(t)
    for r in range(8):
        perm = [j ^ r for j in range(8)]
        sw[..., r, :, :] = t[..., r, perm, :]
    wt = sw.reshape(-1).contiguous().to(device)
    return PackedConvTC(wt, oscale.to(torch.float32).to(device), None if b is None else b.to(torch.float32).contiguous().to(device),
                        cin_p, cout, kh, tile, stride)


def split_f16(x: torch.Tensor, channels: Optional[int] = None, ld: Optional[int] = None) -> Split:
    """fp32 NHWC (B,H,W,ldx) -> Split planes (channel stride ld, default round_up(C, 8))."""
    B, H, W, ldx = x.shape
    C = channels or ldx
    ld = ld or _rup(C, 8)
    hi = torch.empty((B, H, W, ld), device=x.device, dtype=torch.float16)
    lo = torch.empty((B, H, W, ld), device=x.device, dtype=torch.float16)
    if ld != C:
        hi.zero_(); lo.zero_()
    lib().split_f16(ptr(x), B * H * W, C, ldx, ptr(hi), ptr(lo), ld, stream())
    return Split(hi, lo, C)


def conv2d_tc(x: Split, pc: PackedConvTC, res=None, relu: bool = False, out_f32: bool = True, out_split: bool = False,
              nchw: bool = False, out: Optional[torch.Tensor] = None, coff: int = 0, split_ld: Optional[int] = None,
              tapw: Optional[torch.Tensor] = None, r1=None, split_cout: int = 0, extra_ch: int = -1, kernel_select: int = 0):
    """Tensor-core conv.  Returns a dict with the requested outputs: 'y' (fp32 NHWC), 'split' (Split), 'nchw', and with
    ``tapw`` ((9, cout) weights of a following 3x3 -> 1 conv) 'tap': the (B,H,W,12) tap maps contracted in the epilogue."""
    B, Hi, Wi, ldx = x.hi.shape
    pad = pc.k // 2
    H, W = (Hi + 2 * pad - pc.k) // pc.stride + 1, (Wi + 2 * pad - pc.k) // pc.stride + 1
    dev = x.hi.device
    y = None
    if out_f32:
        y = out if out is not None else torch.empty((B, H, W, pc.cout), device=dev, dtype=torch.float32)
    y_nchw = torch.empty((B, pc.cout, H, W), device=dev, dtype=torch.float32) if nchw else None
    sp = None
    if out_split:
        nch = split_cout or pc.cout                      # channels that go to the split planes
        ld = split_ld or _rup(nch, 8)
        if ld != nch:
            sp = Split(torch.zeros((B, H, W, ld), device=dev, dtype=torch.float16),
                       torch.zeros((B, H, W, ld), device=dev, dtype=torch.float16), nch)
        else:
            sp = Split(torch.empty((B, H, W, ld), device=dev, dtype=torch.float16),
                       torch.empty((B, H, W, ld), device=dev, dtype=torch.float16), nch)
    res_f = res if torch.is_tensor(res) else None
    res_s = res if isinstance(res, Split) else None
    tap = torch.empty((B, H, W, 12), device=dev, dtype=torch.float32) if tapw is not None else None
    extra = torch.empty((B, H, W), device=dev, dtype=torch.float32) if extra_ch >= 0 else None
    r1_score, r1_w, r1_bias = r1 if r1 is not None else (None, None, None)      # 65th input channel as a rank-1 term
    lib().conv2d_tc(ptr(x.hi), ptr(x.lo), B, Hi, Wi, pc.cin, ldx, ptr(pc.wt), ptr(pc.oscale), pc.bn, ptr(pc.bias),
                    ptr(res_f), 0 if res_f is None else res_f.shape[3],
                    None if res_s is None else ptr(res_s.hi), None if res_s is None else ptr(res_s.lo),
                    0 if res_s is None else res_s.hi.shape[3],
                    ptr(y), 0 if y is None else y.shape[3], coff, ptr(y_nchw),
                    None if sp is None else ptr(sp.hi), None if sp is None else ptr(sp.lo), 0 if sp is None else sp.hi.shape[3], 0,
                    int(split_cout), ptr(tapw), ptr(tap), ptr(r1_score), ptr(r1_w), ptr(r1_bias), ptr(extra), int(extra_ch),
                    pc.cout, pc.k, pc.k, pc.stride, 1 if relu else 0, int(kernel_select), stream())
    return dict(y=y, split=sp, nchw=y_nchw, tap=tap, extra=extra)


@dataclass
class PackedConv65:
    """3x3 conv whose input is cat(64 channels, score): tensor-core part + fp32 rank-1 part (see frtm_rank1_finish)."""
    main: PackedConvTC
    wx: torch.Tensor            # (9, cout) fp32
    bias: Optional[torch.Tensor]
    cout: int


def pack_conv65(weight: torch.Tensor, bias: Optional[torch.Tensor], device=None) -> PackedConv65:
    cout = weight.shape[0]
    assert weight.shape[1] == 65 and weight.shape[2] == 3
    main = pack_conv_tc(weight[:, :64], None, device=device)
    wx = weight[:, 64].detach().float().permute(1, 2, 0).reshape(9, cout).contiguous().to(device)
    return PackedConv65(main, wx, None if bias is None else bias.detach().float().contiguous().to(device), cout)


def conv65(h: Split, score: torch.Tensor, pc: PackedConv65, n_obj: int = 1, relu: bool = True, want_f32: bool = False,
           want_split: bool = True):
    """h: Split with 64 channels for F = B/n_obj frames (n_obj > 1: shared by the objects of a frame); score (B,H,W) fp32.
    Returns (y fp32 or None, Split(64) or None, extra (B,H,W) fp32 or None)."""
    F, H, W, _ = h.hi.shape
    B = score.shape[0]
    assert B == F * n_obj
    dev = score.device
    if n_obj == 1 and want_split and not want_f32:
        # per-object input: the score channel's rank-1 term, bias and ReLU run in the tensor-core conv's epilogue, which
        # writes the 64-channel split planes and (Cout == 65) the next conv's score channel directly
        o = conv2d_tc(h, pc.main, relu=relu, out_f32=False, out_split=True, split_ld=64, split_cout=64,
                      r1=(score, pc.wx, pc.bias), extra_ch=64 if pc.cout == 65 else -1)
        return None, o["split"], o["extra"]
    ld = _rup(pc.cout, 4)
    main = conv2d_tc(h, pc.main, out=torch.empty((F, H, W, ld), device=dev, dtype=torch.float32))["y"]
    y = torch.empty((B, H, W, pc.cout), device=dev, dtype=torch.float32) if want_f32 else None
    sp = None
    if want_split:
        sp = Split(torch.empty((B, H, W, 64), device=dev, dtype=torch.float16),
                   torch.empty((B, H, W, 64), device=dev, dtype=torch.float16), 64)
    extra = torch.empty((B, H, W), device=dev, dtype=torch.float32) if pc.cout == 65 else None
    lib().rank1_finish(ptr(main), ld, n_obj, ptr(y), pc.cout, ptr(score), ptr(pc.wx), ptr(pc.bias), B, H, W, pc.cout,
                       1 if relu else 0, None if sp is None else ptr(sp.hi), None if sp is None else ptr(sp.lo), 64, ptr(extra),
                       stream())
    return y, sp, extra


def pack_conv_tc_1x1_device(weight: torch.Tensor, bn_tile: Optional[int] = None, out: Optional[PackedConvTC] = None) -> PackedConvTC:
    """(Cout,Cin[,1,1]) fp32 ON THE DEVICE -> PackedConvTC without a host round trip (no synchronisation).  ``out``: a
    previous result of the same shape whose buffers are rewritten in place (stable addresses for captured graphs)."""
    w = weight.detach().reshape(weight.shape[0], -1).contiguous()
    cout, cin = w.shape
    tile = bn_tile or _pick_bn(cout)
    cout_p = (cout + tile - 1) // tile * tile
    if out is not None and out.cin == cin and out.cout == cout and out.bn == tile and out.wt.device == w.device:
        lib().pack_tc_1x1(ptr(w), cout, cin, tile, ptr(out.wt), ptr(out.oscale), stream())
        return out
    wt = torch.empty(cout_p * cin * 2, device=w.device, dtype=torch.float16)
    osc = torch.empty(cout_p, device=w.device, dtype=torch.float32)
    lib().pack_tc_1x1(ptr(w), cout, cin, tile, ptr(wt), ptr(osc), stream())
    return PackedConvTC(wt, osc, None, cin, cout, 1, tile, 1)


def fill_u8(dst: torch.Tensor, vals):
    """Write up to 16 host bytes into a uint8 device tensor asynchronously (values travel as kernel arguments)."""
    import ctypes
    assert dst.dtype == torch.uint8 and dst.numel() >= len(vals)
    for o in range(0, max(len(vals), 1), 16):
        chunk = [int(v) for v in vals[o:o + 16]]
        arr = (ctypes.c_int * max(len(chunk), 1))(*chunk)
        lib().fill_u8(dst.data_ptr() + o, arr, len(chunk), stream())


def fill_small(fdst: Optional[torch.Tensor] = None, fvals=(), idst: Optional[torch.Tensor] = None, ivals=()):
    """Write a few host scalars into device tensors asynchronously (values travel as kernel arguments)."""
    import ctypes
    fa = (ctypes.c_float * max(len(fvals), 1))(*[float(v) for v in fvals])
    ia = (ctypes.c_int * max(len(ivals), 1))(*[int(v) for v in ivals])
    lib().fill_small(ptr(fdst), fa, len(fvals), ptr(idst), ia, len(ivals), stream())
