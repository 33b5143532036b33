"""Refinement ("segmentation") network on libfrtm_b200 kernels.

Drop-in for the reference's ``SegNetwork`` (``model/seg_network.py:149-189``): same constructor, same parameter
names (so reference checkpoints load with ``load_state_dict``), ``forward(scores, features, image_size)`` ->
``(B,1,H,W)`` logits.  Per pyramid level (deep -> shallow): TSE (``:7-21``) -> RRB (``:44-56``) -> CAB (``:24-41``)
-> RRB, then the back-compat upsampler (``:129-146``) built on the fixed x2 bicubic pyramid filter (``:75-126``).

What differs from the reference's execution (not from its results):
  * activations are NHWC and the whole batch of objects (and frames) goes through every kernel at once, instead of
    a Python loop with batch 1 per object (``model/tracker.py:199-204``);
  * ``TSE.reduce(ft)`` depends only on the backbone features, so it is evaluated once per frame and broadcast to
    the objects of that frame (SURVEY.md finding 5);
  * eval-mode BatchNorm is folded into the preceding conv; bias / residual / ReLU are fused conv epilogues;
  * every conv is a tcgen05 tensor-core tile on split-fp16 operands (see ``csrc/conv_tc.cu``); the 65-channel TSE
    convs run as a 64-channel tensor-core conv plus an fp32 rank-1 term for the score channel, so ``cat(h, score)`` is
    never materialised, and the 64-channel part of ``transform[0]`` is shared by all objects of a frame as well.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch
from torch import nn

from .. import ops
from .feature_extractor import FeatureMaps


def _conv(ic, oc, k, bias=True):
    return nn.Conv2d(ic, oc, k, padding=k // 2, bias=bias)


def _relu():
    return nn.LeakyReLU(0.0)


class _Params(nn.Module):
    """Parameter container; the arithmetic lives in SegNetwork.forward."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only")


class TSE(_Params):
    def __init__(self, fc, ic, oc):
        super().__init__()
        nc = ic + oc
        self.reduce = nn.Sequential(_conv(fc, oc, 1), _relu(), _conv(oc, oc, 1))
        self.transform = nn.Sequential(_conv(nc, nc, 3), _relu(), _conv(nc, nc, 3), _relu(), _conv(nc, oc, 3), _relu())


class CAB(_Params):
    def __init__(self, oc, deepest):
        super().__init__()
        self.convreluconv = nn.Sequential(_conv(2 * oc, oc, 1), _relu(), _conv(oc, oc, 1))
        self.deepest = deepest


class RRB(_Params):
    def __init__(self, oc, use_bn=False):
        super().__init__()
        self.conv1x1 = _conv(oc, oc, 1)
        if use_bn:
            self.bblock = nn.Sequential(_conv(oc, oc, 3), nn.BatchNorm2d(oc), _relu(), _conv(oc, oc, 3, bias=False))
        else:
            self.bblock = nn.Sequential(_conv(oc, oc, 3), _relu(), _conv(oc, oc, 3, bias=False))


class BackwardCompatibleUpsampler(_Params):
    def __init__(self, in_channels=64):
        super().__init__()
        self.conv1 = _conv(in_channels, in_channels // 2, 3)
        self.conv2 = _conv(in_channels // 2, 1, 3)


class Upsampler(_Params):
    """The all-frames YouTubeVOS variant's upsampler (``ytvos_validation/seg_network.py:62-74``): true bicubic x2 -> conv1 +
    ReLU -> true bicubic to the image size -> conv2.  Same parameter names as the back-compat module."""

    def __init__(self, in_channels=64):
        super().__init__()
        self.conv1 = _conv(in_channels, in_channels // 2, 3)
        self.conv2 = _conv(in_channels // 2, 1, 3)


class SegNetwork(nn.Module):

    def __init__(self, in_channels=1, out_channels=32, ft_channels=None, use_bn=False, upsampler="backcompat"):
        """``upsampler``: "backcompat" = ``BackwardCompatibleUpsampler`` (``model/seg_network.py:129-146``, the released
        checkpoints); "bicubic" = ``Upsampler`` of ``ytvos_validation/seg_network.py:62-74``."""
        super().__init__()
        if upsampler not in ("backcompat", "bicubic"):
            raise ValueError("upsampler must be 'backcompat' or 'bicubic'")
        self.upsampler = upsampler
        assert ft_channels is not None
        if in_channels != 1:
            raise ValueError("the fused TSE kernels expect a single score channel (in_channels=1)")
        self.ft_channels = ft_channels
        self.use_bn = use_bn
        self.oc = out_channels
        self.TSE, self.RRB1, self.CAB, self.RRB2 = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        for i, (L, fc) in enumerate(self.ft_channels.items()):
            self.TSE[L] = TSE(fc, in_channels, out_channels)
            self.RRB1[L] = RRB(out_channels, use_bn=use_bn)
            self.CAB[L] = CAB(out_channels, i == 0)   # reference: L == 'layer5', i.e. the first (deepest) level
            self.RRB2[L] = RRB(out_channels, use_bn=use_bn)
        self.project = BackwardCompatibleUpsampler(out_channels) if upsampler == "backcompat" else Upsampler(out_channels)
        self._packed = None
        self._bufs: Dict[tuple, torch.Tensor] = {}

    # -- weight packing -----------------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._packed = None
        self._bufs = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def refresh(self):
        """Re-pack weights after an in-place parameter edit."""
        self._packed = None

    def _pack(self):
        dev = self.project.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("SegNetwork must live on a CUDA device (no CPU path)")
        if self.oc != 64:
            raise ValueError("the tensor-core TSE path is specialised for 64 refinement channels")
        P = {}
        oc = self.oc

        def tc(m, **kw):
            return ops.pack_conv_tc(m.weight, m.bias, device=dev, **kw)

        for L in self.ft_channels:
            t, r1, c, r2 = self.TSE[L], self.RRB1[L], self.CAB[L], self.RRB2[L]
            P[L] = dict(
                red0=tc(t.reduce[0]), red2=tc(t.reduce[2]),
                tr0=ops.pack_conv65(t.transform[0].weight, t.transform[0].bias, device=dev),
                tr2=ops.pack_conv65(t.transform[2].weight, t.transform[2].bias, device=dev),
                tr4=ops.pack_conv65(t.transform[4].weight, t.transform[4].bias, device=dev),
                cab_w1=c.convreluconv[0].weight.detach().reshape(oc, 2 * oc).contiguous(), cab_b1=c.convreluconv[0].bias.detach(),
                cab_w2=c.convreluconv[2].weight.detach().reshape(oc, oc).contiguous(), cab_b2=c.convreluconv[2].bias.detach(),
            )
            for tag, r in (("rrb1", r1), ("rrb2", r2)):
                P[L][tag + "_1x1"] = tc(r.conv1x1)
                if self.use_bn:
                    bn = r.bblock[1]
                    P[L][tag + "_a"] = ops.pack_conv_tc(r.bblock[0].weight, r.bblock[0].bias, device=dev, eps=bn.eps, bn=dict(
                        weight=bn.weight.detach(), bias=bn.bias.detach(), running_mean=bn.running_mean, running_var=bn.running_var))
                    P[L][tag + "_b"] = tc(r.bblock[3])
                else:
                    P[L][tag + "_a"] = tc(r.bblock[0])
                    P[L][tag + "_b"] = tc(r.bblock[2])
        P["up1"] = tc(self.project.conv1)
        w2 = self.project.conv2.weight.detach()
        P["up2_w"] = w2.permute(2, 3, 1, 0).reshape(9, w2.shape[1]).contiguous()
        P["up2_b"] = self.project.conv2.bias.detach().contiguous()
        self._packed = P

    # -- forward ------------------------------------------------------------------------------------------------
    def _rrb(self, x_split, W, tag):
        """conv1x1 -> (3x3 + BN + ReLU -> 3x3) + skip -> ReLU, all tensor-core tiles; returns fp32 NHWC."""
        hh = ops.conv2d_tc(x_split, W[tag + "_1x1"], out_f32=False, out_split=True)["split"]
        b = ops.conv2d_tc(hh, W[tag + "_a"], relu=True, out_f32=False, out_split=True)["split"]
        return ops.conv2d_tc(b, W[tag + "_b"], res=hh, relu=True)["y"]

    def forward_nhwc(self, scores: torch.Tensor, feats, image_size) -> torch.Tensor:
        """scores (B,hs,ws) + backbone features as ``ops.Split`` planes (F,h,w,C) with B = F * objects -> logits (B,H,W)."""
        if self._packed is None:
            self._pack()
        P = self._packed
        B = scores.shape[0]
        s_nhwc = scores.reshape(B, scores.shape[-2], scores.shape[-1], 1)
        x = None
        hpool = None
        for li, L in enumerate(self.ft_channels):
            ft = feats[L]
            if not isinstance(ft, ops.Split):
                ft = ops.split_f16(ft)
            F, h, w, _ = ft.hi.shape
            n_obj = B // F
            assert n_obj * F == B, "scores batch must be a multiple of the feature batch"
            W = P[L]
            # TSE.reduce depends on the backbone features only: once per frame, not per object (SURVEY finding 5)
            r = ops.conv2d_tc(ft, W["red0"], relu=True, out_f32=False, out_split=True)["split"]
            o = ops.conv2d_tc(r, W["red2"], out_f32=(li == 0), out_split=True)
            hsp = o["split"]
            if li == 0:
                hpool = ops.global_avgpool(o["y"])                       # (F, oc)
                if n_obj > 1:
                    hpool = hpool.repeat_interleave(n_obj, dim=0)        # tiny (F,64) index plumbing
            s = s_nhwc.reshape(B, h, w) if (h, w) == tuple(s_nhwc.shape[1:3]) else ops.resize_bilinear(s_nhwc, (h, w)).reshape(B, h, w)
            # TSE.transform on cat(h, s): the 64-channel part of the first conv is also shared by the objects of a frame
            _, t, e = ops.conv65(hsp, s, W["tr0"], n_obj=n_obj)
            _, t, e = ops.conv65(t, e, W["tr2"])
            _, t, _ = ops.conv65(t, e, W["tr4"])
            t = self._rrb(t, W, "rrb1")
            # the two global average pools and the gate: two launches (ops.cab with the pools left to it)
            if li == 0:
                t = ops.cab(t, None, hpool, hpool, W["cab_w1"], W["cab_b1"], W["cab_w2"], W["cab_b2"], out_split=True)
            else:
                # F.interpolate(deeper, (h, w)) of seg_network.py:39 happens inside the CAB kernel, which writes the split
                # planes RRB2's first conv reads
                t = ops.cab(t, None, None, x, W["cab_w1"], W["cab_b1"], W["cab_w2"], W["cab_b2"], out_split=True)
            x = self._rrb(t, W, "rrb2")
        # conv2 is linear and so is the bicubic/bilinear chain in front of it: the 32 channels are contracted to the 9 tap
        # maps of conv2 at 240x428 in conv1's epilogue (the 32-channel tensor is never written); one kernel then does
        # bicubic x2 -> bilinear -> sum of the 9 shifted maps.  The x2 in front of conv1 writes conv1's input planes directly.
        if self.upsampler == "bicubic":
            # ytvos_validation Upsampler: the same contraction of conv2 before the (linear) second resize, with true bicubic
            # interpolation (A = -0.75, align_corners=False) in both places
            h, w = x.shape[1:3]
            u = ops.split_f16(ops.resize_bicubic(x, (2 * h, 2 * w)))
            t12 = ops.conv2d_tc(u, P["up1"], relu=True, out_f32=False, tapw=P["up2_w"])["tap"]
            return ops.shift_sum9(ops.resize_bicubic(t12, image_size[-2:]), P["up2_b"])
        u = ops.pyrup_bicubic(x, split=True)
        t12 = ops.conv2d_tc(u, P["up1"], relu=True, out_f32=False, tapw=P["up2_w"])["tap"]
        return ops.upsample_tapsum(t12, P["up2_b"], image_size[-2:])

    def forward(self, scores, features, image_size):
        """Reference signature: scores (B,1,h,w), features dict (NCHW; ``FeatureMaps`` carries NHWC too)."""
        if isinstance(features, FeatureMaps) and all(L in features.split for L in self.ft_channels):
            feats = features.split
        else:
            feats = {L: ops.nchw_to_nhwc(features[L]) for L in self.ft_channels}
        image_size = [int(v) for v in (image_size.tolist() if torch.is_tensor(image_size) else image_size)]
        logits = self.forward_nhwc(scores.reshape(scores.shape[0], *scores.shape[-2:]).contiguous(), feats, image_size)
        return logits.unsqueeze(1)
