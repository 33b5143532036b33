"""ResNet-18/101 backbone feature pass on libfrtm_b200 kernels.

Drop-in for the reference's ``ResnetFeatureExtractor`` (``model/feature_extractor.py:7-87``): same constructor
argument, ``.to(device)``, ``__call__(uint8 image, output_layers) -> {'layer1'..'layer5': (B,C,h,w) fp32}``,
``.get_out_channels()`` (deep→shallow) and ``.no_grad_forward``.  The arithmetic the reference delegates to
torchvision's ``ResNet`` modules (``resnet.py:92-105,146-163``: conv → BN → ReLU, residual add) runs here as NHWC
implicit-GEMM convolutions with eval-mode BatchNorm folded into the weights and bias / residual / ReLU fused in the
epilogue.  The returned dict additionally carries the NHWC working tensors in ``.nhwc`` so the refinement network
consumes them without a layout change.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch

from .. import ops
from .._lib import require_cuda

_ARCH = {"resnet18": ("basic", (2, 2, 2, 2), (64, 64, 128, 256, 512)),
         "resnet101": ("bottleneck", (3, 4, 23, 3), (64, 256, 512, 1024, 2048))}


class FeatureMaps(dict):
    """``dict`` of NCHW feature maps (reference contract) + the working tensors the other kernels consume directly:
    ``.split`` — fp16 hi/lo NHWC planes (operands of the tensor-core convs), ``.nhwc`` — fp32 NHWC tensors of the layers a
    caller of ``forward_split`` asked for through ``f32_layers`` (empty for ``__call__``)."""

    def __init__(self, nchw: Dict[str, torch.Tensor], nhwc: Dict[str, torch.Tensor], split=None):
        super().__init__(nchw)
        self.nhwc = nhwc
        self.split = split or {}


class ResnetFeatureExtractor:

    def __init__(self, name: str = "resnet101", state_dict: Optional[Dict[str, torch.Tensor]] = None):
        if name not in _ARCH:
            raise ValueError("unsupported backbone '%s' (resnet18 | resnet101)" % name)
        self.name = name
        self.kind, self.depth, chans = _ARCH[name]
        if state_dict is None:
            # Same source of weights as the reference (feature_extractor.py:14): torchvision's ImageNet checkpoint.
            import torchvision
            state_dict = getattr(torchvision.models, name)(weights="IMAGENET1K_V1").state_dict()
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items() if not k.startswith("fc.")}
        self._out_channels = OrderedDict(layer5=chans[4], layer4=chans[3], layer3=chans[2], layer2=chans[1], layer1=chans[0])
        self.device = None
        self._w: Dict[str, ops.PackedConv] = {}

    # ------------------------------------------------------------------------------------------------------
    def _bn(self, key):
        return {k: self._sd[key + "." + k] for k in ("weight", "bias", "running_mean", "running_var")}

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("frtm_vos_b200 runs on CUDA (sm_100a) only; got device '%s'" % device)
        self.device = device
        sd, w = self._sd, {}
        # the 7x7/s2 stem as a 1x1 tensor-core conv over the im2col patches (K = 7 x 3 x 8 = 168 -> 192)
        w["stem"] = ops.pack_conv_tc(ops.stem_weight_as_1x1(sd["conv1.weight"]), bn=self._bn("bn1"), device=device)

        def tc(conv_key, bn_key, stride=1):
            return ops.pack_conv_tc(sd[conv_key + ".weight"], bn=self._bn(bn_key), device=device, stride=stride)

        for si, nblk in enumerate(self.depth):
            for bi in range(nblk):
                key = "layer%d.%d" % (si + 1, bi)
                stride = 2 if (bi == 0 and si > 0) else 1
                if self.kind == "basic":
                    w[key + ".c1"] = tc(key + ".conv1", key + ".bn1", stride)
                    w[key + ".c2"] = tc(key + ".conv2", key + ".bn2")
                else:
                    w[key + ".c1"] = tc(key + ".conv1", key + ".bn1")
                    w[key + ".c2"] = tc(key + ".conv2", key + ".bn2", stride)     # torchvision puts the stride on the 3x3
                    w[key + ".c3"] = tc(key + ".conv3", key + ".bn3")
                if (key + ".downsample.0.weight") in sd:
                    w[key + ".ds"] = tc(key + ".downsample.0", key + ".downsample.1", stride)
        self._w = w
        return self

    def get_out_channels(self):
        return self._out_channels

    # ------------------------------------------------------------------------------------------------------
    def forward_split(self, images: torch.Tensor, nchw_layers: Sequence[str] = (), f32_layers: Sequence[str] = (),
                      upto: str = "layer5"):
        """uint8 (B,3,H,W) -> ({layer: Split}, {layer: fp32 NHWC for f32_layers}, {layer: NCHW for nchw_layers}).

        Every conv is a tcgen05 tile with BatchNorm folded in, the residual read from the split planes and ReLU fused:
        1x1 and 3x3, stride 1 and 2, directly; the 7x7/s2 stem as a 1x1 conv over its im2col patches (ops.stem_patches)."""
        if self.device is None:
            raise RuntimeError("call .to(device) first")
        require_cuda(images, "image")
        w = self._w
        split, f32, nchw = {}, {}, {}
        x = ops.stem_conv(images, w["stem"])            # 7x7/s2 stem + bn1 + relu: patches built in shared memory
        if "layer1" in nchw_layers:
            x, nchw["layer1"] = ops.maxpool3x3s2(x, nchw=True)
            xs = ops.split_f16(x)
        else:                                           # the pooling writes the split planes the first block reads
            xs, x = ops.maxpool3x3s2_split(x, want_f32="layer1" in f32_layers)
        if "layer1" in f32_layers:
            f32["layer1"] = x
        split["layer1"] = xs
        last = int(upto[-1])
        for si, nblk in enumerate(self.depth):
            name = "layer%d" % (si + 2)
            if si + 2 > last:
                break
            for bi in range(nblk):
                key = "layer%d.%d" % (si + 1, bi)
                final = bi == nblk - 1
                idt = ops.conv2d_tc(xs, w[key + ".ds"], out_f32=False, out_split=True)["split"] if (key + ".ds") in w else xs
                h = ops.conv2d_tc(xs, w[key + ".c1"], relu=True, out_f32=False, out_split=True)["split"]
                if self.kind == "bottleneck":
                    h = ops.conv2d_tc(h, w[key + ".c2"], relu=True, out_f32=False, out_split=True)["split"]
                last_conv = w[key + (".c2" if self.kind == "basic" else ".c3")]
                o = ops.conv2d_tc(h, last_conv, res=idt, relu=True, out_split=True, out_f32=final and name in f32_layers,
                                  nchw=final and name in nchw_layers)
                xs = o["split"]
                if final:
                    if o["y"] is not None:
                        f32[name] = o["y"]
                    if o["nchw"] is not None:
                        nchw[name] = o["nchw"]
            split[name] = xs
        return split, f32, nchw

    def forward_nhwc(self, images: torch.Tensor, nchw_layers: Sequence[str] = (), upto: str = "layer5"):
        """Compatibility helper: ({layer: fp32 NHWC}, {layer: NCHW})."""
        names = ["layer%d" % i for i in range(1, int(upto[-1]) + 1)]
        split, f32, nchw = self.forward_split(images, nchw_layers, names, upto)
        return f32, nchw

    def __call__(self, input: torch.Tensor, output_layers=None) -> FeatureMaps:
        img = input if input.dim() == 4 else input.unsqueeze(0)
        names = ["layer1", "layer2", "layer3", "layer4", "layer5"]
        want = [n for n in names if output_layers is None or n in output_layers]
        split, f32, nchw = self.forward_split(img, want)   # like the reference, every stage runs regardless of the request
        return FeatureMaps(OrderedDict((n, nchw[n]) for n in want), f32, split)

    def no_grad_forward(self, input, output_layers=None, chunk_size=None):
        if chunk_size is None:
            return self(input, output_layers)
        outs = [self(t, output_layers) for t in torch.split(input, chunk_size)]
        return {L: torch.cat([o[L] for o in outs]) for L in outs[0]}
