"""First-frame augmentation (host logic).

Produces the K=5 training images/masks the target model is initialised on: the original frame plus
four synthetic views in which the object is cut out, the hole is Telea-inpainted, and background and
object are warped independently (rotation / scale / skew / flip / directional blur) and re-composited.
Behaviour follows the reference's ``ImageAugmenter`` (``model/augmenter.py:95-555``) including its
draw order from ``numpy.random`` so that, with the same seed, the same views come out:

* the spec generator always draws ``AugmentationParams2.num_aug - 1 = 19`` specs per round
  (``augmenter.py:44,199``), renders all of them, keeps the valid ones and, if more than ``num_aug-1``
  survive, shuffles and crops (``:516-545``);
* parameter lists are tiled, shuffled in attribute order and sliced (``:204-211``);
* transforms compose ``T(loc)·skew·rot·scale·T(-centre)`` (``:262-265``).

Inpainting stays on the host with OpenCV, exactly like the reference (``:297-340``).  For CPU inputs the
warps run with OpenCV (the reference's CPU path, ``lib/image.py:46-50``) and reproduce the reference
bit for bit; for CUDA inputs the selected views are rendered by libfrtm_b200 kernels (``csrc/augment.cu``:
bicubic affine warp, directional blur, alpha paste) — the B200-native counterpart of the reference's NPP
extension ``lib/_npp/nppig.cpp``.  Like NPP's, the device bicubic is not
bit-identical to OpenCV's fixed-point one; masks (the training labels) always come from the host path.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch
import torch.nn.functional as F

_SPEC_FIELDS = ("location", "rotation", "fliplr", "scale", "skew", "blur_size", "blur_angle")
_SPEC_DEFAULT_POOL = dict(
    num_aug=20,
    location=[(0.5, 0.5)],
    rotation=[5, -5, 10, -10, 20, -20, 30, -30, 45, -45, 60, -60],
    fliplr=[False, False, True],
    scale=[0.7, 1.0, 1.5, 2.0, "0.25", "0.5", "1.0"],
    skew=[(0.0, 0.0), (0.0, 0.0), (0.1, 0.1)],
    blur_size=[0.0, 0.0, 0.0, 2.0, 5.0],
    blur_angle=[0, 45, 90, 135],
)
_SPEC_DEFAULT = dict(location=None, rotation=0.0, fliplr=False, scale=1.0, skew=(0, 0), blur_size=0, blur_angle=0,
                     min_size=10)


def _mat_translate(dx, dy):
    return np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])


def _mat_scale(sx, sy):
    return np.array([[sx, 0, 0], [0, sy, 0], [0, 0, 1]])


def _mat_rotate(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])


def _mat_skew(kx, ky):
    return np.array([[1, kx, 0], [ky, 1, 0], [0, 0, 1]])


def directional_blur_kernel(sx: float, sy: float, rot2: np.ndarray) -> np.ndarray:
    """Rotated anisotropic Gaussian (augmenter.py:120-138)."""
    cov = rot2 @ np.diag((sx, sy)) @ rot2.T
    half = int(np.max((sx, sy)) / 2 + 0.5)
    half = half + (half + 1) % 2
    r = np.arange(-half, half + 1)
    grid = np.stack(np.meshgrid(r, r))
    quad = (grid * np.tensordot(np.linalg.inv(cov), grid, axes=[1, 0])).sum(0)
    g = np.exp(-0.5 * quad)
    return (g / g.sum() * 1.0).astype(np.float32)


def draw_specs(pool: dict, rng=np.random) -> List[dict]:
    """Tile/shuffle/slice every parameter list, then zip into per-view specs (augmenter.py:194-222)."""
    merged = dict(_SPEC_DEFAULT_POOL)
    for k, v in pool.items():
        merged[k] = v            # existing keys keep their position, new keys append (vars() order)
    n = merged["num_aug"] - 1
    cols = {}
    for key, vals in merged.items():
        if key == "num_aug":
            continue
        vals = list(vals) * ((n + len(vals) - 1) // len(vals))
        rng.shuffle(vals)
        cols[key] = vals[:n]
    specs = []
    for i in range(n):
        s = dict(_SPEC_DEFAULT)
        s.update({k: cols[k][i] for k in cols})
        assert s["location"] is not None
        specs.append(s)
    return specs


def spec_transform(spec: dict, bbox, im_size, limit_scale: bool = True, with_blur: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Affine 3x3 + blur kernel of one spec (augmenter.py:224-276).  ``with_blur=False`` skips building the kernel (mask warps
    use the affine part only) and returns ``None`` for it."""
    cx, cy, bw, bh = bbox
    assert bw > 0 and bh > 0
    ih, iw = im_size
    s = spec["scale"]
    if isinstance(s, str):
        s = float(s) * ih / bh
    if limit_scale:
        if s * bw > iw or s * bh > ih:
            s = min(iw / bw, ih / bh)
        if s * bw < spec["min_size"] or s * bh < spec["min_size"]:
            s = max(spec["min_size"] / bw, spec["min_size"] / bh)
    mirror = -1 if spec["fliplr"] else 1
    d2r = np.pi / 180
    loc = spec["location"]
    T = _mat_translate(loc[0] * iw, loc[1] * ih) @ _mat_skew(*spec["skew"]) @ _mat_rotate(spec["rotation"] * d2r) @ \
        _mat_scale(mirror * s, s) @ _mat_translate(-cx, -cy)
    if not with_blur:
        G = None
    elif spec["blur_size"] > 0:
        G = directional_blur_kernel(spec["blur_size"], 0.1, _mat_rotate(spec["blur_angle"] * d2r)[:2, :2])
    else:
        G = np.array([[1.0]], dtype=np.float32)
    return T, G


def warp_affine_host(src: torch.Tensor, H: np.ndarray, size, mode: str = "bicubic") -> torch.Tensor:
    """Per-channel cv2.warpAffine (lib/image.py:38-59, CPU branch)."""
    flat = src.reshape(-1, *src.shape[-2:])
    dst = flat.new_zeros(flat.shape[0], *size)
    H2 = H.astype(np.float32)[:2, :]
    flag = dict(nearest=cv2.INTER_NEAREST, bilinear=cv2.INTER_LINEAR, bicubic=cv2.INTER_CUBIC)[mode]
    for c in range(flat.shape[0]):
        cv2.warpAffine(flat[c].numpy(), H2, (size[1], size[0]), dst[c].numpy(), flag)
    return dst[0] if src.dim() == 2 else dst


def _blur_channels(img: torch.Tensor, kernel: np.ndarray) -> torch.Tensor:
    kernel = np.array(kernel, dtype=np.float32)
    if kernel.shape == (1, 1):
        return img
    fh, fw = kernel.shape
    k = torch.as_tensor(kernel).float().view(1, 1, fh, fw)
    return F.conv2d(img.unsqueeze(1), k, padding=(fh // 2, fw // 2)).squeeze(1)


def _hole_box(hole2d: np.ndarray, radius: int, margin: int, shape):
    """Bounding box (y0, y1, x0, x1) of the nonzero pixels of ``hole2d`` grown by ``margin + 2*radius``, clipped; None if empty."""
    ys = np.flatnonzero(hole2d.any(axis=1))
    xs = np.flatnonzero(hole2d.any(axis=0))
    if ys.size == 0:
        return None
    m = margin + 2 * radius
    return (max(int(ys[0]) - m, 0), min(int(ys[-1]) + m + 1, shape[0]), max(int(xs[0]) - m, 0), min(int(xs[-1]) + m + 1, shape[1]))


def telea_inpaint_cropped(image: np.ndarray, hole: np.ndarray, radius: int, margin: int = 16) -> np.ndarray:
    """``cv2.inpaint(image, hole, radius, INPAINT_TELEA)`` evaluated on the hole's bounding box grown by ``margin`` pixels and
    pasted back.  Telea's fast-marching fill only looks ``radius`` pixels beyond the hole, so the result is identical
    (tests/test_host.py) while the cost no longer scales with the frame."""
    hole2d = hole.reshape(hole.shape[0], hole.shape[1])
    box = _hole_box(hole2d, radius, margin, image.shape[:2])
    if box is None:
        return image.copy()
    y0, y1, x0, x1 = box
    out = image.copy()
    out[y0:y1, x0:x1] = cv2.inpaint(np.ascontiguousarray(image[y0:y1, x0:x1]), np.ascontiguousarray(hole2d[y0:y1, x0:x1]),
                                    inpaintRadius=radius, flags=cv2.INPAINT_TELEA)
    return out


def cut_and_inpaint(im: torch.Tensor, mask: torch.Tensor, d: int = 1, f: int = 1):
    """Object cut-out (RGBA, feathered alpha) + Telea-inpainted background (augmenter.py:297-340)."""
    if d == 1 and f == 1:
        # The configuration the tracker uses (augmenter.py:497).  With 1x1 structuring elements and 1x1 box blurs,
        # erode/blur are identities: alpha = 255*m, and the "blur the inpainted border" blend x*r + (1-r)*x with
        # r in {0,1} returns x exactly, so only the cut-out, the 2x2 dilation and the Telea inpaint remain.  Everything
        # stays planar (C,H,W); only the hole's bounding box is interleaved for cv2.inpaint and pasted back.
        chw = im.detach().cpu().numpy()
        m2 = (mask.squeeze() > 0).byte().detach().cpu().numpy()
        cut = np.empty((4,) + m2.shape, dtype=np.uint8)
        np.multiply(chw, m2[None], out=cut[:3])
        np.multiply(m2, 255, out=cut[3])
        outer = cv2.dilate(m2, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2)))
        bg = chw.copy()
        box = _hole_box(outer, 1, 16, m2.shape)
        if box is not None:
            y0, y1, x0, x1 = box
            patch = cv2.inpaint(np.ascontiguousarray(chw[:, y0:y1, x0:x1].transpose((1, 2, 0))),
                                np.ascontiguousarray(outer[y0:y1, x0:x1]), inpaintRadius=1, flags=cv2.INPAINT_TELEA)
            bg[:, y0:y1, x0:x1] = patch.transpose((2, 0, 1))
        return torch.from_numpy(cut), torch.from_numpy(bg)
    image = im.detach().cpu().numpy().transpose((1, 2, 0))
    m = (mask.squeeze() > 0).byte().detach().cpu().numpy()[..., None]
    cut = m * image
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (f, f))
    alpha = cv2.blur(cv2.erode(m, se) * 255, (f, f))[..., None]
    cut = np.concatenate((cut, alpha), axis=-1)
    inner = cv2.erode(m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (d, d)))
    outer = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (d * 2, d * 2)))
    image = cv2.inpaint(image, outer, inpaintRadius=d, flags=cv2.INPAINT_TELEA)
    ring = cv2.blur((1 - inner) * outer * 255, ksize=(d, d))[..., None] / 255
    soft = cv2.blur(image, ksize=(d, d))
    image = (soft * ring + (1 - ring) * image).astype(np.uint8)
    cut_t = torch.from_numpy(np.ascontiguousarray(cut.transpose((2, 0, 1))))
    bg_t = torch.from_numpy(np.ascontiguousarray(image.transpose((2, 0, 1))))
    return cut_t, bg_t


def mask_center_bbox(mask: torch.Tensor):
    """(cx, cy, w, h) of the nonzero region (augmenter.py:432-452)."""
    m = mask.squeeze()
    ys = m.sum(dim=-1).nonzero(as_tuple=False).view(-1).cpu().numpy()
    xs = m.sum(dim=-2).nonzero(as_tuple=False).view(-1).cpu().numpy()
    if len(ys) > 0 and len(xs) > 0:
        x, y = xs[0], ys[0]
        w, h = xs[-1] - xs[0] + 1, ys[-1] - ys[0] + 1
    else:
        x, y, w, h = 0, 0, 0, 0
    return x + w / 2, y + h / 2, w, h


def target_locations(n: int, im_size, rng=np.random) -> List[Tuple[float, float]]:
    """Jittered grid of new object centres, shuffled (augmenter.py:169-192)."""
    h, w = im_size
    aspect = w / h
    nrows = int(np.ceil(np.sqrt(n / aspect)))
    ncols = int(np.ceil(aspect * nrows))
    pts = []
    for r in range(nrows):
        for c in range(ncols):
            x = (c + 0.5) / ncols + rng.normal(0, 0.5 / ncols / 4)
            y = (r + 0.5) / nrows + rng.normal(0, 0.5 / nrows / 4)
            pts.append((np.round(x, 3), np.round(y, 3)))
    rng.shuffle(pts)
    return pts[:n]


_INPAINT_POOL = None
_BLUR_CACHE = {}


def _inpaint_pool():
    global _INPAINT_POOL
    if _INPAINT_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _INPAINT_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="frtm-inpaint")
    return _INPAINT_POOL


class ImageAugmenter:
    """Drop-in for the reference class of the same name: ``augment_first_frame(im, mask)``."""

    def __init__(self, parameters: dict, device_render: bool = True):
        self.params = parameters
        self.max_retries = 100
        self.device_render = device_render     # CUDA inputs: render the selected views with libfrtm_b200 kernels

    def _warp_mask(self, mask, fg_spec, bbox, size):
        T, _ = spec_transform(fg_spec, bbox, size, with_blur=False)
        return warp_affine_host(mask, np.array(T, dtype=np.float32), size, "nearest")

    def _render(self, bg, cut, fg_spec, bbox, bg_spec):
        """One composited view (augmenter.py:398-430): warp+blur the inpainted background, warp+blur the RGBA cut-out,
        alpha-paste.  The warped mask is produced separately by ``_warp_mask``."""
        size = tuple(bg.shape[-2:])
        if bg_spec is not None:
            h, w = size
            T, G = spec_transform(bg_spec, (w / 2, h / 2, w, h), size, limit_scale=False)
            canvas = warp_affine_host(bg.float(), np.array(T, dtype=np.float32), size).clamp(0, 255)
            canvas = _blur_channels(canvas, G)
        else:
            canvas = bg
        T, G = spec_transform(fg_spec, bbox, size)
        T = np.array(T, dtype=np.float32)
        canvas = canvas.float()
        obj = warp_affine_host(cut.float(), T, size).clamp(0, 255)
        obj = _blur_channels(obj, G)
        a = obj[3].unsqueeze(0) / 255
        return (obj[:3] * a + canvas * (1 - a)).byte()

    def _warp_masks_device(self, mask_d, fg_specs, bbox, size):
        """All candidate masks of a round warped on the GPU (bit-identical to cv2 INTER_NEAREST) + their pixel counts;
        one device->host read of the counts for the whole round."""
        import ctypes
        from .._lib import lib, ptr, stream
        L = lib()
        H, W = size
        n = len(fg_specs)
        out = torch.empty((n, 1, H, W), device=mask_d.device, dtype=torch.uint8)
        counts = torch.zeros(n, device=mask_d.device, dtype=torch.int32)
        src = mask_d.reshape(H, W).contiguous()
        # one launch for the whole round: the n transforms travel as kernel arguments
        Ms = np.stack([np.asarray(spec_transform(fs, bbox, size, with_blur=False)[0], dtype=np.float32)[:2, :] for fs in fg_specs])
        M = (ctypes.c_double * (6 * n))(*Ms.astype(np.float64).ravel())
        L.warp_mask_nearest_batch(ptr(src), H, W, ptr(out), H, W, n, M, 1, ptr(counts), stream())
        return out, counts.tolist()

    def _render_device(self, bg_d, cut_d, fg_spec, bbox, bg_spec):
        """Same view as ``_render`` rendered by libfrtm_b200 kernels: bg_d (3,H,W) / cut_d (4,H,W) uint8 on the GPU."""
        import ctypes
        from .._lib import lib, ptr, stream
        L = lib()
        H, W = bg_d.shape[-2:]
        dev = bg_d.device

        def warp(src, T):
            C = src.shape[0]
            out = torch.empty((C, H, W), device=dev, dtype=torch.float32)
            M = (ctypes.c_double * 6)(*np.asarray(T, dtype=np.float32)[:2, :].astype(np.float64).ravel())
            L.warp_affine(ptr(src), 1, C, H, W, ptr(out), None, H, W, M, 0, 0.0, 255.0, stream())
            return out

        def blur(img, G):
            G = np.asarray(G, dtype=np.float32)
            if G.shape == (1, 1):
                return img
            # the blur kernels of a spec pool are few and recur for every object (the generators are reseeded per object,
            # tracker.py:178-180): keep them on the device instead of a synchronous pageable upload per use
            key = (str(dev), G.shape, G.tobytes())
            k = _BLUR_CACHE.get(key)
            if k is None:
                if len(_BLUR_CACHE) > 256:
                    _BLUR_CACHE.clear()
                k = _BLUR_CACHE[key] = torch.from_numpy(np.ascontiguousarray(G)).to(dev)
            out = torch.empty_like(img)
            L.filter2d(ptr(img), img.shape[0], H, W, ptr(k), G.shape[0], G.shape[1], ptr(out), stream())
            return out

        if bg_spec is not None:
            T, G = spec_transform(bg_spec, (W / 2, H / 2, W, H), (H, W), limit_scale=False)
            canvas = blur(warp(bg_d, T), G)
        else:
            canvas = bg_d.float()
        T, G = spec_transform(fg_spec, bbox, (H, W))
        obj = blur(warp(cut_d, T), G)
        out = torch.empty((3, H, W), device=dev, dtype=torch.uint8)
        L.alpha_paste(ptr(obj), ptr(canvas), H, W, ptr(out), stream())
        return out

    def augment_first_frame(self, im: torch.Tensor, lb: torch.Tensor, rng=None):
        """(3,H,W) u8 + (1,H,W) u8 mask -> ((K,3,H,W) u8, (K,1,H,W) u8) on ``im.device`` (augmenter.py:473-555)."""
        rng = np.random if rng is None else rng     # np.random.RandomState(seed) draws the same stream as seed(seed)
        p = self.params
        dev = im.device
        im_h, lb_h = im.detach().cpu(), lb.detach().cpu()
        size = tuple(im_h.shape[-2:])
        count = int(lb_h.sum())
        no_bg = count == lb_h.numel()
        if count < p["min_px_count"]:
            raise ValueError("Augmentation failed: Target object is too small.")
        bbox = mask_center_bbox(lb_h)
        if tuple(bbox[-2:]) == (0, 0):
            raise ValueError("Augmentation failed: No object to augment.")
        on_device = dev.type == "cuda" and self.device_render
        if on_device:
            # the Telea inpaint (5-20 ms of OpenCV, GIL released) draws no random numbers and is needed only by the
            # rendering at the end: it runs beside the spec drawing / candidate mask warps below
            pending = _inpaint_pool().submit(cut_and_inpaint, im_h, lb_h, 1, 1)
        else:
            cut, bg = cut_and_inpaint(im_h, lb_h, d=1, f=1)

        fg_pool = copy.deepcopy(dict(p["fg_aug_params"]))
        fg_pool["location"] = target_locations(p["num_aug"], size, rng)
        bg_pool = copy.deepcopy(dict(p["bg_aug_params"])) if "bg_aug_params" in p else None
        want = p["num_aug"] - 1
        lo, hi = p["min_px_count"], lb_h.shape[-1] * lb_h.shape[-2] - p["min_px_count"]

        # The reference renders all 19 candidate views of a round and keeps 4 of the valid ones.  Validity depends
        # only on the (cheap, nearest-neighbour) warped mask, so decide first and render only the survivors: same
        # random draws, same selected views, ~5x less host work.
        cand, masks = [], []
        while len(cand) < want:
            fg_specs = draw_specs(fg_pool, rng)
            if bg_pool is not None:
                bg_specs = draw_specs(bg_pool, rng)
            else:
                # the reference pairs the foreground specs with ``[None] * (num_aug - 1)`` and ``zip`` stops there
                # (augmenter.py:520-531): only the first num_aug - 1 candidates of a round are ever looked at
                fg_specs = fg_specs[:want]
                bg_specs = [None] * want
            if on_device:
                warped, counts = self._warp_masks_device(lb.to(dev), fg_specs, bbox, size)
            for j, (fs, bs) in enumerate(zip(fg_specs, bg_specs)):
                if on_device:
                    m, px = warped[j], counts[j]
                else:
                    m = self._warp_mask(lb_h, fs, bbox, size)
                    px = int((m == 1).sum())
                if px >= lo and (px < hi or no_bg):
                    cand.append((fs, bs))
                    masks.append(m)
        if len(cand) > want:
            order = list(range(len(cand)))
            rng.shuffle(order)
            order = order[:want]
            cand = [cand[i] for i in order]
            masks = [masks[i] for i in order]
        if on_device:
            # rendering (bicubic warps, blur, alpha paste) and the nearest-neighbour mask warps on the GPU; inpainting and
            # spec drawing stay on the host like in the reference
            cut, bg = pending.result()
            bg_d, cut_d = bg.to(dev), cut.to(dev)
            views = [self._render_device(bg_d, cut_d, fs, bbox, bs) for fs, bs in cand]
            return torch.stack([im] + views), torch.stack([lb.to(dev).reshape(1, *size)] + masks)
        views = [self._render(bg, cut, fs, bbox, bs) for fs, bs in cand]
        views.insert(0, im_h)
        masks.insert(0, lb_h)
        return torch.stack(views).to(dev), torch.stack(masks).to(dev)
