"""Seeded inputs shared by ``oracle/make_golden.py`` (which runs the REAL reference on them, in the build
container) and by the tests (which run the oracle restatement and the CUDA path on them).  Everything is
drawn from explicit CPU ``torch.Generator`` objects so the tensors are identical on every machine."""
from __future__ import annotations

from collections import OrderedDict

import torch

from frtm_vos_b200 import synth

SMALL = (64, 112)      # image size of the small fixtures -> layer4 map 4x7
MID = (128, 224)       # -> layer4 map 8x14

AUG_PARAMS = dict(
    num_aug=5, min_px_count=1,
    fg_aug_params=dict(rotation=[5, -5, 10, -10, 20, -20, 30, -30, 45, -45], fliplr=[False, False, False, False, True],
                       scale=[0.5, 0.7, 1.0, 1.5, 2.0, 2.5], skew=[(0.0, 0.0), (0.0, 0.0), (0.1, 0.1)],
                       blur_size=[0.0, 0.0, 0.0, 2.0], blur_angle=[0, 45, 90, 135]),
    bg_aug_params=dict(tcenter=[(0.5, 0.5)], rotation=[0, 0, 0], fliplr=[False], scale=[1.0, 1.0, 1.2],
                       skew=[(0.0, 0.0)], blur_size=[0.0, 0.0, 1.0, 2.0, 5.0], blur_angle=[0, 45, 90, 135]),
)


def disc_params(in_channels=256, init_iters=(5, 10), update_iters=(5,), memory_size=20, device="cpu"):
    return dict(layer="layer4", in_channels=in_channels, c_channels=96, out_channels=1, init_iters=tuple(init_iters),
                update_iters=tuple(update_iters), memory_size=memory_size, train_skipping=8, learning_rate=0.1,
                pixel_weighting=dict(method="hinge", tf=0.1), filter_reg=(1e-4, 1e-2), precond=(1e-4, 1e-2),
                precond_lr=0.1, CG_forgetting_rate=750, device=device, update_filters=True)


def oracle_disc_params(dp: dict) -> dict:
    """Reference-style disc_params -> kwargs of ``oracle.frtm_ref.TargetModelRef``."""
    d = {k: v for k, v in dp.items() if k not in ("out_channels", "pixel_weighting", "update_filters", "device")}
    d["tf"] = dp["pixel_weighting"]["tf"]
    return d


def blob_masks(n, size, gen, soft=False):
    """n smooth blob masks (n,1,H,W) in [0,1] (soft) or {0,1}."""
    H, W = size
    yy = torch.arange(H).float().view(1, H, 1)
    xx = torch.arange(W).float().view(1, 1, W)
    cy = (0.25 + 0.5 * torch.rand(n, 1, 1, generator=gen)) * H
    cx = (0.25 + 0.5 * torch.rand(n, 1, 1, generator=gen)) * W
    ry = (0.12 + 0.2 * torch.rand(n, 1, 1, generator=gen)) * H
    rx = (0.12 + 0.2 * torch.rand(n, 1, 1, generator=gen)) * W
    d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
    m = torch.sigmoid((1.0 - d) * 4.0)
    if not soft:
        m = (m > 0.5).float()
    return m.unsqueeze(1)


def update_problem(seed=5, M=12, cap=16, c=96, fsize=(4, 7), isize=SMALL):
    """A filled frame memory for the update-phase (filter-only) GN problem."""
    g = torch.Generator().manual_seed(seed)
    samples = torch.zeros(cap, c, *fsize)
    samples[:M] = torch.randn(M, c, *fsize, generator=g) * 0.5
    labels = torch.zeros(cap, 1, *isize)
    labels[:M] = blob_masks(M, isize, g, soft=True)
    pw = torch.zeros(cap, 1, *isize)
    pw[:M] = 0.5 + torch.rand(M, 1, *isize, generator=g)
    w = torch.zeros(cap)
    w[:M] = torch.rand(M, generator=g) + 0.1
    w /= w.sum()
    F0 = torch.randn(1, c, 3, 3, generator=g) * 0.05
    return dict(samples=samples, labels=labels, pixel_weights=pw, weights=w, F0=F0)


def init_problem(seed=6, K=5, C=256, c=96, fsize=(4, 7), isize=SMALL):
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(K, C, *fsize, generator=g))
    y = blob_masks(K, isize, g, soft=False)
    P0 = torch.randn(c, C, 1, 1, generator=g) * (1.0 / C ** 0.5)
    F0 = torch.randn(1, c, 3, 3, generator=g) * 0.05
    dP = torch.randn(c, C, 1, 1, generator=g) * 0.1
    dF = torch.randn(1, c, 3, 3, generator=g) * 0.1
    return dict(x=x, y=y, P0=P0, F0=F0, dP=dP, dF=dF)


def feedforward_case(arch="resnet18", size=MID, n_obj=2, seed=9):
    """Fixed-state feed-forward case: image + per-object (P, F) + network weights."""
    seq = synth.SyntheticSequence(num_objects=n_obj, num_frames=4, size=size, seq_id=seed)
    bb = synth.backbone_state_dict(arch, size=size)
    seg = synth.segnet_state_dict(arch)
    C = synth.backbone_out_channels(arch)["layer4"]
    g = torch.Generator().manual_seed(seed)
    PF = []
    for _ in range(n_obj):
        PF.append((torch.randn(96, C, 1, 1, generator=g) * (1.0 / C ** 0.5), torch.randn(1, 96, 3, 3, generator=g) * 0.03))
    return dict(image=seq[2][0], bb=bb, seg=seg, PF=PF, arch=arch)


def ytvos_case(seed=41):
    """Seeded inputs of the all-frames variant's own pieces: an ``Upsampler`` state + input, object
    probabilities for ``merge_segmentations``, a soft mask for the 'thresh' update labels / hinge weights."""
    g = torch.Generator().manual_seed(seed)
    up = OrderedDict()
    up["project.conv1.weight"] = torch.randn(32, 64, 3, 3, generator=g) / 24.0
    up["project.conv1.bias"] = torch.randn(32, generator=g) * 0.1
    up["project.conv2.weight"] = torch.randn(1, 32, 3, 3, generator=g) / 17.0
    up["project.conv2.bias"] = torch.randn(1, generator=g) * 0.1
    x = torch.randn(2, 64, 9, 13, generator=g)
    probs = torch.rand(3, 4, 20, 30, generator=g)               # (objects, frames, H, W)
    probs[1, 2] = 0.0
    probs[:, 3, :5] = 1.0
    soft = blob_masks(3, (48, 64), g, soft=True)
    soft[2] = soft[2] * 0.01                                     # fewer than 10 pixels above 0.5
    return dict(up=up, x=x, image_size=(70, 101), probs=probs, soft=soft)


def strip_prefix(sd, prefix="refiner."):
    return OrderedDict(((k[len(prefix):] if k.startswith(prefix) else k), v) for k, v in sd.items())


# ---- evaluation: seeded label maps for the J / F measures -------------------------------------------
EVAL_SIZES = [(48, 64), (37, 53), (96, 128), (8, 8)]


def eval_mask_pairs(seed=31):
    """[(annotation, segmentation)] boolean maps: blobs and a shifted / eroded / noisy copy, plus the degenerate cases the
    measures special-case (both empty, one empty, full frame, single pixel, touching the border)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    pairs = []
    for (h, w) in EVAL_SIZES:
        yy, xx = np.mgrid[:h, :w]
        for _ in range(4):
            cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
            ry, rx = rng.uniform(0.1, 0.35) * h, rng.uniform(0.1, 0.35) * w
            gt = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
            dy, dx = rng.randint(-3, 4), rng.randint(-3, 4)
            sg = np.roll(np.roll(gt, dy, 0), dx, 1) ^ (rng.uniform(size=(h, w)) > 0.97)
            pairs.append((gt, sg))
        z, o = np.zeros((h, w), bool), np.ones((h, w), bool)
        one = z.copy(); one[h // 2, w // 2] = True
        edge = z.copy(); edge[:3, :] = True; edge[:, -2:] = True
        pairs += [(z, z), (z, pairs[-1][0]), (pairs[-1][0], z), (o, o), (o, one), (one, one), (edge, np.roll(edge, 1, 0))]
    return pairs


def eval_score_vectors(seed=32):
    """Per-frame score vectors with NaNs (frames before the start / the last frame) for the sequence statistics."""
    import numpy as np
    rng = np.random.RandomState(seed)
    out = []
    for n in (4, 5, 9, 24, 70, 104):
        v = rng.uniform(0, 1, size=n)
        v[0] = np.nan
        v[-1] = np.nan
        if n > 8:
            v[1:rng.randint(1, n // 2)] = np.nan
        out.append(v)
    out.append(np.array([np.nan, 0.2, 0.9, 1.0, 0.0, 0.51, 0.5, np.nan]))
    return out
