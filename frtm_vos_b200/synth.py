"""Seeded synthetic inputs for the FRTM hot path.

Nothing here is on the compute path: these are *data generators* used by
``bench.py``, ``__graft_entry__.smoke()`` and the tests so that the CUDA path,
the oracle and (in the build container) the real reference all read the same
tensors.  Three generators:

* :class:`SyntheticSequence` — smoothly textured, colour-tinted moving objects over a
  rolling smooth background, implementing the sequence protocol the reference driver consumes
  (``lib/datasets.py:16-69``: ``name``, ``obj_ids``, ``frame_names``, ``len``,
  ``[i] -> (image u8 (3,H,W), labels u8 (1,H,W) | [], new_obj_ids)``,
  ``preload(device)``).
* :func:`backbone_state_dict` — torchvision-layout ResNet-18/101 weights,
  residual-damped (last BN of every block γ=0.25) and BN-calibrated on synthetic
  frames so deep activations stay O(1) (plain random init explodes on rn101).
* :func:`segnet_state_dict` — structured pass-through checkpoint for the
  refinement network (keys as in ``evaluate.py:144`` / ``model/seg_network.py``)
  so that masks are meaningful and the online update actually fires.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

_PALETTE = torch.tensor([[0.9, 0.2, 0.2], [0.2, 0.8, 0.3], [0.2, 0.3, 0.9], [0.9, 0.8, 0.2], [0.8, 0.2, 0.8], [0.2, 0.8, 0.8]])
BACKBONE_BLOCKS = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet101": ("bottleneck", (3, 4, 23, 3))}


# --------------------------------------------------------------------------------------
# video
# --------------------------------------------------------------------------------------
class SyntheticSequence:
    """N textured objects (rectangles / ellipses) translating over a smooth background."""

    def __init__(self, num_objects: int = 3, num_frames: int = 65, size: Tuple[int, int] = (480, 854),
                 seq_id: int = 0, name: Optional[str] = None, start_frames: Optional[Sequence[int]] = None):
        self.H, self.W = int(size[0]), int(size[1])
        self.num_frames = int(num_frames)
        self.name = name or ("synth%03d" % seq_id)
        self.dset_name = "synthetic"
        self.obj_ids = list(range(1, num_objects + 1))
        self.frame_names = ["%05d" % i for i in range(self.num_frames)]
        self.seq_id = seq_id
        self._start = list(start_frames) if start_frames is not None else [0] * num_objects
        assert len(self._start) == num_objects
        self._device = None
        self._frames: Optional[List[torch.Tensor]] = None
        self._build()

    # -- construction ---------------------------------------------------------------
    def _build(self):
        H, W, N = self.H, self.W, len(self.obj_ids)
        g = torch.Generator().manual_seed(1000 * self.seq_id + 999)
        gh, gw = max(H // 16, 2), max(W // 16 + 1, 2)
        coarse = torch.rand(1, 3, gh, gw, generator=g)
        self._bg = F.interpolate(coarse, (H, W), mode="bilinear", align_corners=False)[0]  # (3,H,W) in [0,1)

        # start layout: one grid cell per object, object centred in its cell
        cols = int(math.ceil(math.sqrt(N * W / H)))
        rows = int(math.ceil(N / cols))
        cw, ch = W / cols, H / rows
        self._objs = []
        for k in range(N):
            gk = torch.Generator().manual_seed(1000 * self.seq_id + k)
            area = (0.03 + 0.09 * torch.rand(1, generator=gk).item()) * H * W
            aspect = 0.6 + 0.8 * torch.rand(1, generator=gk).item()
            oh = min(math.sqrt(area / aspect), 0.8 * ch)
            ow = min(area / max(oh, 1.0), 0.8 * cw)
            oh, ow = max(int(oh), 6), max(int(ow), 6)
            cx = (k % cols + 0.5) * cw
            cy = (k // cols + 0.5) * ch
            vel = (torch.rand(2, generator=gk) * 6.0 - 3.0).tolist()
            ellipse = bool(k % 2)
            # smooth texture (noise on a 16-px grid) tinted with a per-object colour: a random-init backbone is not
            # shift-invariant at stride 16, so per-pixel noise textures do not survive a 2-3 px motion (measured:
            # score IoU 0.95 -> 0.30 one frame after init); this recipe keeps IoU 0.87-0.95 over 16 frames.
            coarse_t = torch.rand(1, 3, max(oh // 16, 2), max(ow // 16, 2), generator=gk)
            smooth = F.interpolate(coarse_t, (oh, ow), mode="bilinear", align_corners=False)[0]
            tex = 0.7 * (0.5 * smooth + 0.25) + 0.3 * _PALETTE[k % len(_PALETTE)].view(3, 1, 1)
            yy = (torch.arange(oh).float() + 0.5) / oh * 2 - 1
            xx = (torch.arange(ow).float() + 0.5) / ow * 2 - 1
            shape = ((yy[:, None] ** 2 + xx[None, :] ** 2) <= 1.0) if ellipse else torch.ones(oh, ow, dtype=torch.bool)
            self._objs.append(dict(cx=cx, cy=cy, vx=vel[0], vy=vel[1], h=oh, w=ow, tex=tex, shape=shape))

    @staticmethod
    def _reflect(p: float, lo: float, hi: float) -> float:
        span = hi - lo
        if span <= 0:
            return lo
        q = (p - lo) % (2 * span)
        return lo + (q if q <= span else 2 * span - q)

    def _render(self, t: int) -> Tuple[torch.Tensor, torch.Tensor]:
        H, W = self.H, self.W
        im = torch.roll(self._bg, shifts=2 * t, dims=2).clone()
        lb = torch.zeros(H, W, dtype=torch.uint8)
        for k, o in enumerate(self._objs):
            if t < self._start[k]:
                continue
            x0 = self._reflect(o["cx"] - o["w"] / 2 + o["vx"] * t, 0, W - o["w"])
            y0 = self._reflect(o["cy"] - o["h"] / 2 + o["vy"] * t, 0, H - o["h"])
            x0, y0 = int(round(x0)), int(round(y0))
            sl = (slice(y0, y0 + o["h"]), slice(x0, x0 + o["w"]))
            m = o["shape"]
            im[:, sl[0], sl[1]] = torch.where(m[None], o["tex"], im[:, sl[0], sl[1]])
            lb[sl] = torch.where(m, torch.full_like(lb[sl], k + 1), lb[sl])
        im_u8 = (im * 255.0).clamp(0, 255).to(torch.uint8)
        return im_u8, lb[None]

    # -- sequence protocol ------------------------------------------------------------
    def __len__(self):
        return self.num_frames

    def ground_truth(self, t: int) -> torch.Tensor:
        return self._render(t)[1]

    def __getitem__(self, item: int):
        if item < 0 or item >= self.num_frames:
            raise IndexError(item)
        new_ids = [self.obj_ids[k] for k in range(len(self.obj_ids)) if self._start[k] == item]
        if self._frames is not None:
            im = self._frames[item]
            lb = self.ground_truth(item) if new_ids else []
        else:
            im, lbt = self._render(item)
            lb = lbt if new_ids else []
        if new_ids:
            keep = torch.zeros(256, dtype=torch.bool)
            keep[torch.tensor(new_ids)] = True
            lb = torch.where(keep[lb.long()], lb, torch.zeros_like(lb))
        return im, lb, new_ids

    def preload(self, device):
        """Upload all frames (outside the fps timer, like the reference's ``preload``)."""
        self._frames = [self._render(t)[0].to(device) for t in range(self.num_frames)]
        self._device = device

    def __repr__(self):
        return "synthetic: %s, %d frames, %d objects" % (self.name, self.num_frames, len(self.obj_ids))


# --------------------------------------------------------------------------------------
# backbone weights
# --------------------------------------------------------------------------------------
def backbone_state_dict(name: str = "resnet18", size: Tuple[int, int] = (480, 854), damping: float = 0.25,
                        seed: int = 0, calib_frames: Sequence[int] = (0, 5, 10, 15)) -> "OrderedDict[str, torch.Tensor]":
    """Seeded, residual-damped, BN-calibrated torchvision-layout ResNet weights (no avgpool/fc).

    Data generation only: uses the stock ``torchvision`` module in train mode to accumulate
    cumulative BN statistics over a few synthetic frames, then returns its ``state_dict``.
    """
    import torchvision

    torch.manual_seed(seed)
    net = getattr(torchvision.models, name)(weights=None)
    kind, _ = BACKBONE_BLOCKS[name]
    last_bn = "bn3" if kind == "bottleneck" else "bn2"
    with torch.no_grad():
        for stage in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in stage:
                getattr(blk, last_bn).weight.fill_(damping)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    seq = SyntheticSequence(num_objects=3, num_frames=max(calib_frames) + 1, size=size, seq_id=0)
    frames = torch.stack([seq[t][0] for t in calib_frames]).float()
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 3, 1, 1)
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 3, 1, 1)
    x = frames * (1.0 / 255.0 / std) + (-mean / std)
    net.train()
    with torch.no_grad():
        for _ in range(2):
            h = net.maxpool(net.relu(net.bn1(net.conv1(x))))
            h = net.layer4(net.layer3(net.layer2(net.layer1(h))))
    net.eval()
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items()
                     if not k.startswith("fc.") and not k.endswith("num_batches_tracked"))
    return sd


def backbone_out_channels(name: str) -> "OrderedDict[str, int]":
    """Deep→shallow channel table (``model/feature_extractor.py:20-25``)."""
    if name == "resnet18":
        ch = (512, 256, 128, 64, 64)
    elif name == "resnet101":
        ch = (2048, 1024, 512, 256, 64)
    else:
        raise ValueError("unsupported backbone '%s'" % name)
    return OrderedDict(zip(("layer5", "layer4", "layer3", "layer2", "layer1"), ch))


# --------------------------------------------------------------------------------------
# refinement-network checkpoint
# --------------------------------------------------------------------------------------
def segnet_param_shapes(ft_channels: Dict[str, int], nch: int = 64, in_ch: int = 1, use_bn: bool = True):
    """Names and shapes of the refinement network parameters (checkpoint layout of the reference)."""
    shp: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    nc = nch + in_ch

    def conv(key, oc, ic, k, bias=True):
        shp[key + ".weight"] = (oc, ic, k, k)
        if bias:
            shp[key + ".bias"] = (oc,)

    def bnorm(key, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            shp[key + "." + s] = (c,)
        shp[key + ".num_batches_tracked"] = ()

    for L, fc in ft_channels.items():
        conv("TSE.%s.reduce.0" % L, nch, fc, 1)
        conv("TSE.%s.reduce.2" % L, nch, nch, 1)
        conv("TSE.%s.transform.0" % L, nc, nc, 3)
        conv("TSE.%s.transform.2" % L, nc, nc, 3)
        conv("TSE.%s.transform.4" % L, nch, nc, 3)
    for grp in ("RRB1", "RRB2"):
        for L in ft_channels:
            conv("%s.%s.conv1x1" % (grp, L), nch, nch, 1)
            conv("%s.%s.bblock.0" % (grp, L), nch, nch, 3)
            if use_bn:
                bnorm("%s.%s.bblock.1" % (grp, L), nch)
                conv("%s.%s.bblock.3" % (grp, L), nch, nch, 3, bias=False)
            else:
                conv("%s.%s.bblock.2" % (grp, L), nch, nch, 3, bias=False)
    for L in ft_channels:
        conv("CAB.%s.convreluconv.0" % L, nch, 2 * nch, 1)
        conv("CAB.%s.convreluconv.2" % L, nch, nch, 1)
    conv("project.conv1", nch // 2, nch, 3)
    conv("project.conv2", 1, nch // 2, 3)
    return shp


def segnet_state_dict(backbone: str = "resnet18", seed: int = 7, gain: float = 12.0,
                      prefix: str = "refiner.") -> "OrderedDict[str, torch.Tensor]":
    """Structured pass-through refinement checkpoint: ``logit ≈ gain·(mean_L relu(score) − 0.5)`` + noise."""
    chans = backbone_out_channels(backbone)
    ft_channels = OrderedDict((L, chans[L]) for L in ("layer5", "layer4", "layer3", "layer2"))
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in segnet_param_shapes(ft_channels).items():
        if key.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_var") or (".bblock.1.weight" in key):
            t = torch.ones(shape)
        elif ".bblock.1." in key:
            t = torch.zeros(shape)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (0.02 / math.sqrt(fan_in))
        else:
            t = torch.randn(shape, generator=g) * 0.002
        sd[key] = t
    for L in ft_channels:
        sd["TSE.%s.transform.0.weight" % L][64, 64, 1, 1] += 1.0
        sd["TSE.%s.transform.2.weight" % L][64, 64, 1, 1] += 1.0
        sd["TSE.%s.transform.4.weight" % L][0, 64, 1, 1] += 1.0
        sd["RRB1.%s.conv1x1.weight" % L][0, 0, 0, 0] += 1.0
        sd["RRB2.%s.conv1x1.weight" % L][0, 0, 0, 0] += 1.0
        sd["CAB.%s.convreluconv.2.bias" % L] += 8.0
    sd["project.conv1.weight"][0, 0, 1, 1] += 1.0
    sd["project.conv2.weight"][0, 0, 1, 1] += gain / 4.0
    sd["project.conv2.bias"] -= gain / 2.0
    return OrderedDict((prefix + k, v) for k, v in sd.items())
