"""ORACLE — test infrastructure only, never the product path.

A plain-PyTorch (fp32, CPU or any torch device) restatement of the FRTM per-frame
inference path of the reference (andr345/frtm-vos).  It exists to CHECK the
hand-written CUDA path in ``frtm_vos_b200`` and to serve as the timed CPU baseline in
``bench.py`` (``cpu_baseline`` / ``--impl reference``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline leg of ``bench.py`` may import it.

Pinning: the reference ships no tests / golden vectors, so this
restatement is pinned against the *executed reference itself*: ``oracle/make_golden.py``
imports the real reference from ``/root/reference`` in the build container (with the
non-invasive shims of ``oracle/shims.py``), runs both on the same seeded inputs and
asserts agreement; the resulting fixtures are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle_golden.py`` wherever the repo travels.

Every function cites the reference lines it restates.  The reference's arithmetic on
this path lives partly in third-party code that is not under ``/root/reference``:
torchvision's ResNet (unpinned by the reference, ``README.md:28``; 0.26.0 here) and
ATen's conv / bilinear / autograd kernels; those are called here through the same
``torch.nn.functional`` entry points, in the same order.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# a1 — backbone feature pass  (model/feature_extractor.py:27-68 + torchvision resnet.py)
# --------------------------------------------------------------------------------------
_STAGES = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet101": ("bottleneck", (3, 4, 23, 3))}
_IM_MEAN = (0.485, 0.456, 0.406)
_IM_STD = (0.229, 0.224, 0.225)


def normalize_image(img: Tensor) -> Tensor:
    """uint8 (3,H,W) or (B,3,H,W) -> float (B,3,H,W); model/feature_extractor.py:27-32,42."""
    std = torch.tensor(_IM_STD, dtype=torch.float, device=img.device).reshape(1, 3, 1, 1)
    mean = torch.tensor(_IM_MEAN, dtype=torch.float, device=img.device).reshape(1, 3, 1, 1)
    return (1 / 255 / std) * img.float() + (-mean / std)


def _bn(sd, key, x):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        False, 0.0, 1e-5)


def _basic_block(sd, key, x, stride):
    # torchvision resnet.py:92-105
    idt = x
    h = F.relu(_bn(sd, key + ".bn1", F.conv2d(x, sd[key + ".conv1.weight"], None, stride, 1)))
    h = _bn(sd, key + ".bn2", F.conv2d(h, sd[key + ".conv2.weight"], None, 1, 1))
    if (key + ".downsample.0.weight") in sd:
        idt = _bn(sd, key + ".downsample.1", F.conv2d(x, sd[key + ".downsample.0.weight"], None, stride, 0))
    return F.relu(h + idt)


def _bottleneck(sd, key, x, stride):
    # torchvision resnet.py:146-163 (stride on the 3x3, :135)
    idt = x
    h = F.relu(_bn(sd, key + ".bn1", F.conv2d(x, sd[key + ".conv1.weight"])))
    h = F.relu(_bn(sd, key + ".bn2", F.conv2d(h, sd[key + ".conv2.weight"], None, stride, 1)))
    h = _bn(sd, key + ".bn3", F.conv2d(h, sd[key + ".conv3.weight"]))
    if (key + ".downsample.0.weight") in sd:
        idt = _bn(sd, key + ".downsample.1", F.conv2d(x, sd[key + ".downsample.0.weight"], None, stride, 0))
    return F.relu(h + idt)


def backbone_features(sd: Dict[str, Tensor], arch: str, img: Tensor,
                      output_layers: Optional[Sequence[str]] = None) -> Dict[str, Tensor]:
    """Five named maps layer1..layer5 (model/feature_extractor.py:40-68)."""
    kind, depth = _STAGES[arch]
    block = _basic_block if kind == "basic" else _bottleneck
    out: Dict[str, Tensor] = {}

    def keep(name, t):
        if output_layers is None or name in output_layers:
            out[name] = t

    x = normalize_image(img)
    x = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    keep("layer1", x)
    for si, nblk in enumerate(depth):
        for bi in range(nblk):
            x = block(sd, "layer%d.%d" % (si + 1, bi), x, 2 if (bi == 0 and si > 0) else 1)
        keep("layer%d" % (si + 2), x)
    return out


# --------------------------------------------------------------------------------------
# shared fixed operators (lib/utils.py:25-41)
# --------------------------------------------------------------------------------------
def resize_bilinear(t: Tensor, size) -> Tensor:
    """lib/utils.py:33-35 — skipped when the size already matches."""
    size = tuple(int(s) for s in size)
    if tuple(t.shape[-2:]) == size:
        return t
    return F.interpolate(t, size, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------
# a3 — refinement network (model/seg_network.py:7-189)
# --------------------------------------------------------------------------------------
_CUBIC_E = (-0.10546875, 0.87890625, 0.26171875, -0.03515625)  # Keys a=-0.75 at phase -0.25


def pyr_up_bicubic(x: Tensor) -> Tensor:
    """Fixed x2 bicubic pyramid upsample (model/seg_network.py:75-126)."""
    c = x.shape[1]
    we = torch.tensor(_CUBIC_E, dtype=x.dtype, device=x.device)
    wo = we.flip(0)
    a = F.pad(x, (2, 2, 2, 2), "replicate")
    planes = []
    for wr in (we, wo):          # row phase
        for wc in (we, wo):      # column phase
            k = (wr[:, None] @ wc[None, :]).expand(c, 1, 4, 4).contiguous()
            planes.append(F.conv2d(a, k, groups=c))
    i00, i01, i10, i11 = planes
    n, _, h, w = i11.shape
    top = torch.stack((i00, i01), dim=-1).view(n, c, h, 2 * w)
    bot = torch.stack((i10, i11), dim=-1).view(n, c, h, 2 * w)
    out = torch.stack((top, bot), dim=-2).view(n, c, 2 * h, 2 * w)
    return out[:, :, 1:-1, 1:-1]


def _conv(sd, key, x, k):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), 1, k // 2)


def _rrb(sd, key, x, use_bn):
    # model/seg_network.py:44-56
    h = _conv(sd, key + ".conv1x1", x, 1)
    b = _conv(sd, key + ".bblock.0", h, 3)
    if use_bn:
        b = _bn(sd, key + ".bblock.1", b)
        b = _conv(sd, key + ".bblock.3", F.relu(b), 3)
    else:
        b = _conv(sd, key + ".bblock.2", F.relu(b), 3)
    return F.relu(h + b)


def seg_forward(sd: Dict[str, Tensor], scores: Tensor, features: Dict[str, Tensor], image_size,
                layers: Sequence[str] = ("layer5", "layer4", "layer3", "layer2"), use_bn: bool = True,
                taps: Optional[dict] = None, upsampler: str = "backcompat") -> Tensor:
    """scores (B,1,h,w) + feature maps -> logits (B,1,H,W) (model/seg_network.py:176-189).

    ``sd`` keys are un-prefixed (``TSE.layer4.reduce.0.weight`` …).
    """
    x = None
    for L in layers:
        ft = features[L]
        s = resize_bilinear(scores, ft.shape[-2:])
        # TSE (:7-21)
        h = _conv(sd, "TSE.%s.reduce.2" % L, F.relu(_conv(sd, "TSE.%s.reduce.0" % L, ft, 1)), 1)
        hpool = F.adaptive_avg_pool2d(h, (1, 1)) if x is None else x
        h = torch.cat((h, resize_bilinear(s, h.shape[-2:])), dim=1)
        for j in (0, 2, 4):
            h = F.relu(_conv(sd, "TSE.%s.transform.%d" % (L, j), h, 3))
        h = _rrb(sd, "RRB1.%s" % L, h, use_bn)
        # CAB (:24-41)
        sp = F.adaptive_avg_pool2d(h, (1, 1))
        dp = hpool if L == layers[0] else F.adaptive_avg_pool2d(hpool, (1, 1))
        gate = _conv(sd, "CAB.%s.convreluconv.2" % L,
                     F.relu(_conv(sd, "CAB.%s.convreluconv.0" % L, torch.cat((sp, dp), dim=1), 1)), 1)
        h = h * torch.sigmoid(gate) + resize_bilinear(hpool, h.shape[-2:])
        x = _rrb(sd, "RRB2.%s" % L, h, use_bn)
        if taps is not None:
            taps[L] = x
    if upsampler == "bicubic":
        return upsampler_bicubic(sd, x, image_size)
    # BackwardCompatibleUpsampler (:129-146)
    x = pyr_up_bicubic(x)
    x = F.relu(_conv(sd, "project.conv1", x, 3))
    x = pyr_up_bicubic(x)
    x = F.interpolate(x, tuple(int(v) for v in image_size[-2:]), mode="bilinear", align_corners=False)
    return _conv(sd, "project.conv2", x, 3)


def upsampler_bicubic(sd: Dict[str, Tensor], x: Tensor, image_size) -> Tensor:
    """``Upsampler`` of the all-frames YouTubeVOS variant (ytvos_validation/seg_network.py:62-74)."""
    x = F.interpolate(x, (2 * x.shape[-2], 2 * x.shape[-1]), mode="bicubic", align_corners=False)
    x = F.relu(_conv(sd, "project.conv1", x, 3))
    x = F.interpolate(x, tuple(int(v) for v in image_size[-2:]), mode="bicubic", align_corners=False)
    return _conv(sd, "project.conv2", x, 3)


# --------------------------------------------------------------------------------------
# a5 — hinge pixel weights (model/discriminator.py:107-152)
# --------------------------------------------------------------------------------------
def pixel_weights(y: Tensor, tf: float = 0.1) -> Tensor:
    n, c, hh, ww = y.shape
    y = y.float()
    px = y.sum(dim=(2, 3))
    af = (px / (hh * ww)).view(n, c, 1, 1)
    px = px.view(n, c, 1, 1)
    small = (px < 10).float()
    af = small * tf + (1 - small) * af
    big = (af > tf).float()
    tfe = big * af + (1 - big) * tf
    wf = tfe / af
    wb = (1 - tfe) / (1 - af)
    return torch.sqrt(wf * y + wb * (1 - y))


# --------------------------------------------------------------------------------------
# a6 — frame memory (model/memory.py:4-92)
# --------------------------------------------------------------------------------------
class FrameMemory:
    def __init__(self, capacity: int, feat_shape, label_shape, device, lr: float):
        self.samples = torch.zeros(capacity, *feat_shape, device=device)
        self.weights = torch.zeros(capacity, device=device)
        self.labels = torch.zeros(capacity, *label_shape, device=device)
        self.pixel_weights = torch.zeros(capacity, *label_shape, device=device)
        self.capacity = capacity
        self.size = 0
        self.prev_ind: Optional[int] = None
        self.lr = lr

    def fill(self, feats: Tensor, labels: Tensor, pw: Tensor):
        """memory.py:33-46 — first sample gets double weight."""
        k = feats.shape[0]
        self.samples[:k] = feats.detach()
        self.weights[:k] = 1.0 / k
        self.weights[0] = 2.0 / k
        self.weights[:k] = self.weights[:k] / self.weights[:k].sum()
        self.labels[:k] = labels.float()
        self.pixel_weights[:k] = pw
        self.size = k

    def _next_slot(self) -> int:
        """memory.py:65-92 — replace the (first) minimum-weight slot."""
        sw, lr = self.weights, self.lr
        if self.size == 0 or lr == 1:
            sw[:] = 0
            sw[0] = 1
            r = 0
        else:
            r = int(torch.min(sw, 0)[1].item())
            if self.prev_ind is None:
                sw /= (1 - lr)
                sw[r] = lr
            else:
                sw[r] = sw[self.prev_ind] / (1 - lr)
        sw /= sw.sum()
        return r

    def insert(self, feat: Tensor, label: Tensor, pw: Tensor):
        """memory.py:48-63."""
        self.prev_ind = self._next_slot()
        self.samples[self.prev_ind] = feat.detach()
        self.labels[self.prev_ind] = label
        self.pixel_weights[self.prev_ind] = pw
        self.size = min(self.size + 1, self.capacity)


# --------------------------------------------------------------------------------------
# a7/a8 — Gauss-Newton + Polak-Ribiere CG through autograd, exactly as the reference does
#          (model/discriminator.py:11-64, model/optimizer.py:18-160)
# --------------------------------------------------------------------------------------
class GNProblem:
    """Weighted least squares on the upsampled score map + Tikhonov terms (discriminator.py:38-64)."""

    def __init__(self, memory: FrameMemory, regs: Sequence[float], precond: Sequence[float], project: bool):
        self.mem = memory
        self.regs = list(regs)
        self.precond = list(precond)
        self.project = project
        self.x = self.y = self.w = None

    def gather(self):
        """discriminator.py:38-43 — active samples into fresh tensors, w = pw*sqrt(sw)."""
        a = self.mem.weights > 0.0
        self.x = self.mem.samples[a]
        self.y = self.mem.labels[a]
        self.w = self.mem.pixel_weights[a] * self.mem.weights[a].sqrt().view(-1, 1, 1, 1)

    def residuals(self, theta: List[Tensor]) -> List[Tensor]:
        """discriminator.py:45-50."""
        if self.project:
            s = F.conv2d(F.conv2d(self.x, theta[0]), theta[1], None, 1, 1)
        else:
            s = F.conv2d(self.x, theta[0], None, 1, 1)
        s = F.interpolate(s, self.y.shape[-2:], mode="bilinear", align_corners=False)
        return [self.w * (s - self.y)] + [r * t for r, t in zip(self.regs, theta)]

    @staticmethod
    def dot(a: List[Tensor], b: List[Tensor]) -> Tensor:
        """discriminator.py:52-61 — sum of per-entry dot products (0-dim)."""
        parts = [u.reshape(-1) @ v.reshape(-1) for u, v in zip(a, b)]
        tot = 0
        for p in parts:
            tot = tot + p.unsqueeze(0)
        return tot


class GaussNewtonCGRef:
    """optimizer.py:18-160 with fletcher_reeves=False, standard_alpha=True, step_alpha=1."""

    def __init__(self, problem: GNProblem, theta: List[Tensor], forget: float):
        self.problem = problem
        self.theta = theta            # leaf tensors, updated in place
        self.forget = forget
        self.p: Optional[List[Tensor]] = None
        self.rho = torch.ones(1)
        self.r_prev: Optional[List[Tensor]] = None
        self.trace: Optional[list] = None   # optional per-iteration dump for parity tests

    def run(self, cg_iters: Sequence[int]):
        self.problem.gather()
        for n in cg_iters:
            self._gn_step(int(n))
        for t in self.theta:
            t.detach_()

    def _gn_step(self, n_cg: int):
        """optimizer.py:77-91."""
        th = self.theta
        for t in th:
            t.requires_grad_(True)
        f0 = self.problem.residuals(th)
        g = [f.detach().requires_grad_(True) for f in f0]
        jt_g = list(torch.autograd.grad(f0, th, g, create_graph=True))
        b = [-v.detach() for v in jt_g]

        def apply_A(v: List[Tensor]) -> List[Tensor]:
            """optimizer.py:155-157 — J^T J v by double backward."""
            jv = torch.autograd.grad(jt_g, g, v, retain_graph=True)
            return list(torch.autograd.grad(f0, th, jv, retain_graph=True))

        dx = self._cg(b, apply_A, n_cg)
        for t in th:
            t.detach_()
        for t, d in zip(th, dx):
            t += 1.0 * d

    def _cg(self, b, apply_A, n_iter: int):
        """optimizer.py:98-153."""
        dot = self.problem.dot
        if self.forget == 0:
            self.p, self.rho, self.r_prev = None, torch.ones(1), None
        elif self.p is not None:
            self.rho = self.rho / self.forget
        r = [v.clone() for v in b]
        x = None
        for it in range(n_iter):
            z = [v / m for v, m in zip(r, self.problem.precond)]
            rho1 = self.rho
            self.rho = dot(r, z)
            if self.p is None:
                self.p = [v.clone() for v in z]
            else:
                rho2 = dot(self.r_prev, z)
                beta = ((self.rho - rho2) / rho1).clamp(0)
                self.p = [zz + pp * beta for zz, pp in zip(z, self.p)]
            q = apply_A(self.p)
            pq = dot(self.p, q)
            alpha = self.rho / pq
            self.r_prev = [v.clone() for v in r]
            if x is None:
                x = [pp * alpha for pp in self.p]
            else:
                x = [xx + pp * alpha for xx, pp in zip(x, self.p)]
            if self.trace is not None:
                self.trace.append(dict(p=[v.clone() for v in self.p], q=[v.clone() for v in q],
                                       alpha=alpha.clone(), rho=self.rho.clone()))
            if it < n_iter - 1:
                r = [rr - qq * alpha for rr, qq in zip(r, q)]
        return x


# --------------------------------------------------------------------------------------
# a2/a3(disc) — target model (model/discriminator.py:67-227)
# --------------------------------------------------------------------------------------
class TargetModelRef:
    def __init__(self, in_channels: int, c_channels: int = 96, init_iters=(5, 10, 10, 10, 10), update_iters=(10,),
                 filter_reg=(1e-4, 1e-2), precond=(1e-4, 1e-2), precond_lr: float = 0.1,
                 CG_forgetting_rate: int = 750, memory_size: int = 80, train_skipping: int = 8,
                 learning_rate: float = 0.1, tf: float = 0.1, device="cpu", seed_weights: Optional[Tuple[Tensor, Tensor]] = None):
        self.device = torch.device(device)
        if seed_weights is None:
            # same default init as nn.Conv2d (kaiming_uniform, a=sqrt(5)) drawn in the same order
            pm = torch.nn.Conv2d(in_channels, c_channels, 1, bias=False)
            fm = torch.nn.Conv2d(c_channels, 1, 3, padding=1, bias=False)
            P, Fw = pm.weight.detach().clone(), fm.weight.detach().clone()
        else:
            P, Fw = seed_weights
        self.P = P.to(self.device)
        self.F = Fw.to(self.device)
        self.init_iters, self.update_iters = tuple(init_iters), tuple(update_iters)
        self.filter_reg, self.precond = tuple(filter_reg), tuple(precond)
        self.forget = (1 - precond_lr) ** CG_forgetting_rate
        self.memory_size, self.train_skipping, self.lr, self.tf = memory_size, train_skipping, learning_rate, tf
        self.frame_num = 0
        self.memory: Optional[FrameMemory] = None
        self.optimizer: Optional[GaussNewtonCGRef] = None
        self.current_sample: Optional[Tensor] = None

    def init(self, x: Tensor, y: Tensor):
        """discriminator.py:154-199."""
        pw = pixel_weights(y, self.tf)
        mem0 = FrameMemory(y.shape[0], x.shape[-3:], y.shape[-3:], self.device, self.lr)
        mem0.fill(x, y, pw)
        opt0 = GaussNewtonCGRef(GNProblem(mem0, self.filter_reg, self.precond, True), [self.P, self.F], self.forget)
        opt0.run(self.init_iters)
        cx = F.conv2d(x, self.P)
        mem = FrameMemory(self.memory_size, cx.shape[-3:], y.shape[-3:], self.device, self.lr)
        mem.fill(cx, y, pw)
        opt = GaussNewtonCGRef(GNProblem(mem, self.filter_reg[1:], self.precond[1:], False), [self.F], self.forget)
        opt.run(self.update_iters)
        self.memory, self.optimizer = mem, opt

    def apply(self, ft: Tensor) -> Tensor:
        """discriminator.py:201-206."""
        self.frame_num += 1
        cft = F.conv2d(ft, self.P)
        self.current_sample = cft
        return F.conv2d(cft, self.F, None, 1, 1)

    def update(self, train_y: Tensor, method: str = "soft") -> bool:
        """discriminator.py:208-227; returns True when a GN update ran.  ``method='thresh'``: the all-frames variant stores
        the binarised mask as the training label (ytvos_validation/discriminator.py:364-367)."""
        if self.current_sample is None:
            return False
        if (train_y > 0.5).sum() < 10:
            return False
        ys = (train_y > 0.5).float()
        self.memory.insert(self.current_sample, ys if method == "thresh" else train_y, pixel_weights(ys, self.tf))
        if self.frame_num % self.train_skipping != 0:
            return False
        self.optimizer.run(self.update_iters)
        return True


# --------------------------------------------------------------------------------------
# a4 — multi-object merge + labels (model/tracker.py:143-150, 208-221)
# --------------------------------------------------------------------------------------
def merge_masks(masks: Tensor) -> Tensor:
    """tracker.py:214-221: (N+1,H,W) probabilities -> argmax-gated softmax scores (same shape)."""
    p = torch.clamp(masks, 1e-7, 1 - 1e-7)
    p[0:1] = torch.min((1 - p[1:]), dim=0, keepdim=True)[0]
    segs = F.softmax(p / (1 - p), dim=0)
    inds = segs.argmax(dim=0)
    out = torch.empty_like(masks)
    for i in range(masks.shape[0]):
        out[i] = segs[i] * (inds == i).float()
    return out


def labels_from_masks(masks: Tensor, lut: Tensor, single_object: bool) -> Tensor:
    """tracker.py:143-150."""
    if single_object:
        return lut[(masks[1:2] > 0.5).long()]
    m = torch.clamp(masks, 1e-7, 1 - 1e-7)
    m[0:1] = torch.min((1 - m[1:]), dim=0, keepdim=True)[0]
    segs = F.softmax(m / (1 - m), dim=0)
    return lut[segs.argmax(dim=0)]


# --------------------------------------------------------------------------------------
# a10 + sequence driver (model/tracker.py:37-227)
# --------------------------------------------------------------------------------------
class TrackerRef:
    """Sequence driver.  ``augment(image, mask) -> (K,3,H,W) u8, (K,1,H,W) u8`` is a plug-in, as in the reference."""

    def __init__(self, backbone_sd, arch: str, seg_sd, disc_params: dict, augment, device="cpu",
                 seg_layers=("layer5", "layer4", "layer3", "layer2"), use_bn: bool = True, hooks: Optional[dict] = None):
        self.device = torch.device(device)
        self.arch = arch
        self.bb = {k: v.to(self.device) for k, v in backbone_sd.items()}
        self.seg = {(k[len("refiner."):] if k.startswith("refiner.") else k): v.to(self.device) for k, v in seg_sd.items()}
        self.disc_params = dict(disc_params)
        self.layer = self.disc_params.pop("layer", "layer4")
        self.disc_params.pop("device", None)
        self.augment = augment
        self.seg_layers, self.use_bn = tuple(seg_layers), use_bn
        self.targets: "OrderedDict[int, dict]" = OrderedDict()
        self.current_masks: Optional[Tensor] = None
        self.current_frame = 0
        self.hooks = hooks or {}

    def _features(self, img, layers=None):
        with torch.no_grad():
            return backbone_features(self.bb, self.arch, img, layers)

    def initialize(self, image: Tensor, labels: Tensor, new_objects: Sequence[int]):
        """tracker.py:165-191."""
        import numpy as np
        self.current_masks = torch.zeros((len(self.targets) + len(new_objects) + 1, *image.shape[-2:]), device=self.device)
        for oid in new_objects:
            mask = (labels == oid).byte()
            tm = TargetModelRef(device=self.device, **self.disc_params)
            tgt = dict(id=oid, index=len(self.targets) + 1, start_frame=self.current_frame, start_mask=mask, model=tm)
            self.targets[oid] = tgt
            torch.random.manual_seed(0)
            np.random.seed(0)
            im, msk = self.augment(image, mask)
            ft = self._features(im, [self.layer])
            if "init_inputs" in self.hooks:
                self.hooks["init_inputs"](oid, im, msk, ft[self.layer], tm)
            tm.init(ft[self.layer], msk)
            if "after_init" in self.hooks:
                self.hooks["after_init"](oid, tm)
            self.current_masks[tgt["index"]] = mask
        return self.current_masks

    def track(self, image: Tensor):
        """tracker.py:193-227."""
        im_size = image.shape[-2:]
        feats = self._features(image)
        live = [t for t in self.targets.values() if t["start_frame"] < self.current_frame]
        for t in live:
            with torch.no_grad():
                s = t["model"].apply(feats[self.layer])
                logits = seg_forward(self.seg, s, feats, im_size, self.seg_layers, self.use_bn)
            if "logits" in self.hooks:
                self.hooks["logits"](self.current_frame, t["id"], s, logits)
            self.current_masks[t["index"]] = torch.sigmoid(logits)
        for t1 in live:
            for t2 in self.targets.values():
                if t1["id"] != t2["id"] and t2["start_frame"] == self.current_frame:
                    self.current_masks[t1["index"]] *= (1 - t2["start_mask"].squeeze(0)).float()
        self.current_masks = merge_masks(self.current_masks)
        for t in live:
            ran = t["model"].update(self.current_masks[t["index"]].unsqueeze(0).unsqueeze(0))
            if ran and "after_update" in self.hooks:
                self.hooks["after_update"](self.current_frame, t["id"], t["model"])
        return self.current_masks

    def run_sequence(self, sequence):
        """tracker.py:103-163 — returns (list of uint8 label maps, fps incl. per-object init)."""
        from time import time
        self.targets = OrderedDict()
        self.current_frame = 0
        lut = torch.tensor([0] + list(sequence.obj_ids), dtype=torch.uint8, device=self.device)
        outputs = []
        t0 = time()
        for i in range(len(sequence)):
            image, labels, new_objects = sequence[i]
            had_targets = len(self.targets) > 0
            image = image.to(self.device)
            if len(new_objects) > 0:
                labels = labels.to(self.device)
                self.initialize(image, labels, new_objects)
            if had_targets:
                self.track(image)
                labels = labels_from_masks(self.current_masks, lut, len(sequence.obj_ids) == 1)
            if isinstance(labels, list) and len(labels) == 0:
                labels = image.new_zeros(1, *image.shape[-2:])
            outputs.append(labels)
            self.current_frame += 1
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        dt = time() - t0
        return outputs, len(sequence) / dt


# --------------------------------------------------------------------------------------
# f4 — the all-frames YouTubeVOS variant of the driver (ytvos_validation/tracker.py:53-207)
# --------------------------------------------------------------------------------------
def merge_segmentations(fg: Tensor) -> Tensor:
    """ytvos_validation/tracker.py:53-62 — fg (N, ...) object probabilities -> softmax over {background, objects}."""
    fg = torch.clamp(fg, 1e-7, 1 - 1e-7)
    bg = torch.min((1 - fg), dim=0, keepdim=True)[0]
    p = torch.cat((bg, fg), dim=0)
    return F.softmax(p / (1 - p), dim=0)


class YtvosTrackerRef(TrackerRef):
    """Every object initialised up front from its own first frame, raw probabilities per frame, ground truth re-inserted,
    ONE merge over all frames (``:82-207``); bicubic ``Upsampler``; ``update_method='thresh'``."""

    def __init__(self, *a, update_method: str = "thresh", **k):
        super().__init__(*a, **k)
        self.update_method = update_method

    def _track_frame(self, image: Tensor, index: Dict[int, int], n_obj: int) -> Tensor:
        im_size = image.shape[-2:]
        out = torch.zeros((n_obj, *im_size), device=self.device)
        live = [t for t in self.targets.values() if t["start_frame"] < self.current_frame]
        if not live:
            return out
        feats = self._features(image)
        ys = {}
        for t in live:
            with torch.no_grad():
                s = t["model"].apply(feats[self.layer])
                logits = seg_forward(self.seg, s, feats, im_size, self.seg_layers, self.use_bn, upsampler="bicubic")
            if "logits" in self.hooks:
                self.hooks["logits"](self.current_frame, t["id"], s, logits)
            ys[t["id"]] = torch.sigmoid(logits)[0, 0]
        if self.current_frame > 0:                               # update (:133-161)
            upd = torch.zeros((n_obj, *im_size), device=self.device)
            for t1 in live:
                for t2 in self.targets.values():
                    if t1["id"] != t2["id"] and t2["start_frame"] == self.current_frame:
                        ys[t1["id"]] = ys[t1["id"]] * (1 - t2["start_mask"].reshape(*im_size)).float()
                upd[index[t1["id"]]] = ys[t1["id"]]
            segs = merge_segmentations(upd)
            inds = segs.argmax(dim=0)
            for i in range(n_obj):
                m = inds == (i + 1)
                upd[i] = torch.where(m, segs[i + 1], torch.zeros_like(segs[i + 1]))
            for t in live:
                t["model"].update(upd[index[t["id"]]].unsqueeze(0).unsqueeze(0), method=self.update_method)
        for t in live:
            out[index[t["id"]]] = ys[t["id"]]
        return out

    def run_sequence(self, sequence):
        from time import time
        ids = list(sequence.obj_ids)
        items = [sequence[i] for i in range(len(sequence))]
        first = {}
        for i, (im, lb, new) in enumerate(items):
            for oid in new:
                first[oid] = i
        self.targets = OrderedDict()
        t0 = time()
        for f0 in sorted(set(first.values())):
            self.current_frame = f0
            self.initialize(items[f0][0].to(self.device), items[f0][1].to(self.device), [o for o in ids if first[o] == f0])
        index = {o: k for k, o in enumerate(ids)}
        outs = []
        for i in range(len(items)):
            self.current_frame = i
            outs.append(self._track_frame(items[i][0].to(self.device), index, len(ids)))
        out = torch.stack(outs)                                   # (T, N, H, W)
        for o in ids:
            out[first[o], index[o]] = (items[first[o]][1].to(self.device)[0] == o).float()
        segs = merge_segmentations(out.permute(1, 0, 2, 3))
        lut = torch.tensor([0] + ids, dtype=torch.uint8, device=self.device)
        labels = lut[segs.argmax(dim=0)]                          # (T, H, W)
        return [labels[i].unsqueeze(0) for i in range(len(items))], len(items) / (time() - t0)
