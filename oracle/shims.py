"""ORACLE tooling — import the *real* reference from /root/reference (build container only).

The reference is read-only, GPL, and does not travel to the GPU box; nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may use this module.  It is used by ``oracle/make_golden.py``
(fixture generation) and by the ``needs_reference`` tests that cross-check ``oracle/frtm_ref.py``
against the executed reference.  The reference files are untouched; the shims below only make it
importable/runnable offline on torch 2.11 / numpy 2:

1. ``easydict`` stand-in (attribute dict)                        — evaluate.py:4
2. ``skimage.morphology`` stand-in (``disk`` restated, ``binary_dilation`` = scipy.ndimage's, which is what
   scikit-image calls) and ``numpy1_aliases`` (``np.bool`` / ``np.int``) — lib/davis.py:5,63,160
3. ``lib._npp`` stub (the NPP warp extension JIT-builds into ~/tmp on import and is only ever
   *called* for CUDA tensors)                                    — lib/image.py:6, lib/_npp/__init__.py
4. torchvision ``resnetXX(pretrained=True)`` -> ``weights=None`` — model/feature_extractor.py:12-14
5. ``TensorList.__getattr__`` must reject dunder names           — lib/tensorlist.py:169-176
6. ``torch.cuda.synchronize`` no-op on a CPU-only host           — model/tracker.py:126,159
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FRTM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "tracker.py"))


class _AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            return _AttrDict(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        import copy
        return _AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _disk(radius, dtype="uint8"):
    """scikit-image's published ``morphology.disk``: all grid points within ``radius`` of the centre."""
    import numpy as np
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)


def _binary_dilation(image, footprint=None, out=None):
    """scikit-image's ``morphology.binary_dilation`` is ``scipy.ndimage.binary_dilation`` with the footprint as structure."""
    import scipy.ndimage as ndi
    return ndi.binary_dilation(image, structure=footprint, output=out)


class numpy1_aliases:
    """``with numpy1_aliases():`` restores ``np.bool`` / ``np.int`` (removed in numpy 1.24) while reference code that
    still uses them runs (lib/davis.py:63-64,160, lib/utils.py:16)."""

    def __enter__(self):
        import numpy as np
        self._added = [n for n in ("bool", "int") if n not in np.__dict__]
        for n in self._added:
            setattr(np, n, {"bool": bool, "int": int}[n])
        return self

    def __exit__(self, *exc):
        import numpy as np
        for n in self._added:
            delattr(np, n)
        return False


_loaded = None


def load_reference():
    """Returns a namespace with the reference modules (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch
    import torchvision

    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _AttrDict
        sys.modules["easydict"] = m
    try:
        import skimage.morphology  # noqa: F401
    except Exception:
        sk = types.ModuleType("skimage")
        mo = types.ModuleType("skimage.morphology")
        mo.binary_dilation, mo.disk = _binary_dilation, _disk
        sk.morphology = mo
        sys.modules["skimage"], sys.modules["skimage.morphology"] = sk, mo
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import lib  # the reference's package (ours live under frtm_vos_b200/)
    assert os.path.realpath(os.path.dirname(lib.__file__)).startswith(os.path.realpath(REFERENCE_ROOT))
    npp = types.ModuleType("lib._npp")
    npp.nppig_cpp = None
    sys.modules["lib._npp"] = npp

    import lib.tensorlist as tl
    _orig = tl.TensorList.__getattr__

    def _guarded(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _orig(self, name)

    tl.TensorList.__getattr__ = _guarded

    import model.feature_extractor as fe
    for n in ("resnet18", "resnet34", "resnet50", "resnet101"):
        setattr(fe, n, (lambda ctor: (lambda pretrained=True: ctor(weights=None)))(getattr(torchvision.models, n)))

    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None

    import model.tracker as tracker
    import model.discriminator as discriminator
    import model.optimizer as optimizer
    import model.memory as memory
    import model.seg_network as seg_network
    import model.augmenter as augmenter

    import lib.datasets as datasets
    import lib.davis as davis
    import lib.evaluation as evaluation
    import lib.utils as utils
    import lib.image as image

    ns = types.SimpleNamespace(datasets=datasets, davis=davis, evaluation=evaluation, utils=utils, image=image,
                               numpy1_aliases=numpy1_aliases, tracker=tracker, discriminator=discriminator, optimizer=optimizer, memory=memory,
                               seg_network=seg_network, feature_extractor=fe, tensorlist=tl, augmenter=augmenter,
                               EasyDict=_AttrDict)
    _loaded = ns
    return ns
