"""J / F measures of the DAVIS benchmark as the reference evaluates them (``lib/davis.py:19-236``,
itself an adaptation of the DAVIS-2017 toolkit), written for numpy >= 2 (the reference uses the removed ``np.bool``) and
without scikit-image: the boundary dilation with a disk runs through ``cv2.dilate``.

Definitions kept exactly (checked against the executed reference in ``tests/test_eval_cpu.py``):
* J: intersection over union of two binary maps, 1 when both are empty (``:53-71``).
* F: boundary F-measure — one-pixel boundaries (a pixel differs from its east, south or south-east neighbour,
  ``:137-194``) matched within ``ceil(0.008 * |(H, W)|)`` pixels by dilating with a disk (``:75-134``).
* per object: the measure on every frame strictly between the object's start frame and the last frame, NaN elsewhere
  (``:37-43``); statistics mean / recall(>0.5) / decay(first quarter - last quarter) / std ignoring NaNs (``:197-236``).
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import Dict

import cv2
import numpy as np


def _as_bool(a) -> np.ndarray:
    a = np.asarray(a)
    return a if a.dtype == np.bool_ else a != 0


def davis_jaccard_measure(fg_mask, gt_mask) -> float:
    fg, gt = _as_bool(fg_mask), _as_bool(gt_mask)
    union = int(np.count_nonzero(fg | gt))
    if union == 0:
        return 1
    return np.count_nonzero(fg & gt) / np.float64(union)


def seg2bmap(seg, width=None, height=None) -> np.ndarray:
    """Binary map of one-pixel-wide boundaries, offset by half a pixel towards the origin (``:137-194``).  Only the
    same-size case the evaluation uses is provided."""
    seg = _as_bool(seg)
    assert np.atleast_3d(seg).shape[2] == 1
    h, w = seg.shape[:2]
    if (width is not None and width != w) or (height is not None and height != h):
        raise NotImplementedError("seg2bmap: resampled boundary maps are not used by the evaluation")
    b = np.zeros((h, w), dtype=np.bool_)
    # interior: differs from the east, south or south-east neighbour
    b[:-1, :-1] = (seg[:-1, :-1] ^ seg[:-1, 1:]) | (seg[:-1, :-1] ^ seg[1:, :-1]) | (seg[:-1, :-1] ^ seg[1:, 1:])
    b[-1, :-1] = seg[-1, :-1] ^ seg[-1, 1:]          # last row: east neighbour only
    b[:-1, -1] = seg[:-1, -1] ^ seg[1:, -1]          # last column: south neighbour only
    return b


def _disk(radius: float) -> np.ndarray:
    r = int(np.floor(radius))
    ax = np.arange(-r, r + 1)
    return ((ax[:, None] ** 2 + ax[None, :] ** 2) <= radius * radius).astype(np.uint8)


def _dilate(mask: np.ndarray, selem: np.ndarray) -> np.ndarray:
    # zero outside the image, symmetric structuring element: the same set as a binary dilation with that footprint
    return cv2.dilate(mask.astype(np.uint8), selem, borderType=cv2.BORDER_CONSTANT, borderValue=0) != 0


def davis_f_measure(foreground_mask, gt_mask, bound_th=0.008) -> float:
    fg, gt = _as_bool(foreground_mask), _as_bool(gt_mask)
    assert np.atleast_3d(fg).shape[2] == 1
    bound_pix = bound_th if bound_th >= 1 else np.ceil(bound_th * np.linalg.norm(fg.shape))
    fg_b, gt_b = seg2bmap(fg), seg2bmap(gt)
    n_fg, n_gt = int(np.count_nonzero(fg_b)), int(np.count_nonzero(gt_b))
    if n_fg == 0 and n_gt > 0:
        precision, recall = 1, 0
    elif n_fg > 0 and n_gt == 0:
        precision, recall = 0, 1
    elif n_fg == 0 and n_gt == 0:
        precision, recall = 1, 1
    else:
        selem = _disk(float(bound_pix))
        precision = np.count_nonzero(fg_b & _dilate(gt_b, selem)) / float(n_fg)
        recall = np.count_nonzero(gt_b & _dilate(fg_b, selem)) / float(n_gt)
    if precision + recall == 0:
        return 0
    return 2 * precision * recall / (precision + recall)


def nanmean(*args, **kwargs):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmean(*args, **kwargs)


def mean(X):
    """Average ignoring NaNs (warns on an all-NaN input exactly like ``np.nanmean``)."""
    return np.nanmean(X)


def recall(X, threshold=0.5):
    """Fraction of the non-NaN values above ``threshold``."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        x = np.asarray(X)
        x = x[~np.isnan(x)]
        return mean(x > threshold)


def decay(X, n_bins=4):
    """Mean of the first quarter minus mean of the last quarter of the non-NaN values (``:215-229``).  The reference
    casts its bin edges to uint8; sequences here are shorter than 256 evaluated frames, where that is the identity —
    longer ones use the un-wrapped edges."""
    x = np.asarray(X)
    x = x[~np.isnan(x)]
    ids = (np.round(np.linspace(1, len(x), n_bins + 1) + 1e-10) - 1).astype(np.int64)
    bins = [x[ids[i]:ids[i + 1] + 1] for i in range(0, 4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmean(bins[0]) - np.nanmean(bins[3])


def std(X):
    return np.nanstd(X)


_MEASURES = {"J": davis_jaccard_measure, "F": davis_f_measure}
_STATISTICS = OrderedDict((("decay", decay), ("mean", mean), ("recall", recall), ("std", std)))


def _label_map(t) -> np.ndarray:
    a = t.numpy() if hasattr(t, "numpy") else np.asarray(t)
    return a[0] if a.ndim == 3 else a


def evaluate_sequence(segmentations: Dict, annotations: Dict, object_info: Dict, measure="J") -> dict:
    """``segmentations`` / ``annotations``: ordered {frame name: (1,H,W) label map}; ``object_info``: {object id: name of
    its start frame}.  Returns ``{'raw': {id: per-frame scores}, 'decay': [...], 'mean': [...], 'recall': [...], 'std': [...]}``
    with one entry per object in ``object_info`` order (``:19-49``)."""
    fn = _MEASURES[measure]
    names = list(annotations.keys())
    seg_names = list(segmentations.keys())
    n = len(names)
    results = dict(raw=OrderedDict())
    for obj_id, first_frame in object_info.items():
        r = np.full(n, np.nan)
        first = names.index(first_frame)
        for i in range(first + 1, min(n - 1, len(seg_names))):
            r[i] = fn(_label_map(annotations[names[i]]) == obj_id, _label_map(segmentations[seg_names[i]]) == obj_id)
        results["raw"][obj_id] = r
    for stat, stat_fn in _STATISTICS.items():
        results[stat] = [float(stat_fn(r)) for r in results["raw"].values()]
    return results
