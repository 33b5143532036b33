"""Inference datasets: sequences backed by JPEG frames and first-frame PNG labels.

Same classes, constructor arguments, attributes and item protocol as the reference (``lib/datasets.py:16-158``):
``FileSequence`` (``.name``, ``.obj_ids``, ``.frame_names``, ``.start_frames``, ``len()``, ``[i] -> (uint8 (3,H,W),
uint8 (1,H,W) | [], list[int])``, ``.preload(device)``), ``DAVISDataset`` and ``YouTubeVOSDataset``.

What is different is how the frames get to the GPU.  The reference decodes and uploads one frame at a time on the
calling thread (``lib/datasets.py:63-65``: ``imread(f).to(device)`` per frame, a pageable synchronous copy each).  Here
``preload_async`` decodes on a small thread pool (PIL releases the GIL while decoding) straight into ONE pinned staging
slab per sequence, each frame is sent with an asynchronous copy on a dedicated copy stream the moment it is decoded, and
``Tracker.run_dataset`` starts the preload of sequence k+1 before it tracks sequence k, so JPEG decode and H2D of the
next sequence hide behind the current one.  ``preload(device)`` (the reference call) = start if necessary + wait, so the
fps definition (``model/tracker.py:91,130,159-161``: preload is outside the timer) is unchanged.
"""
from __future__ import annotations

import json
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from .image import imread

_DECODE_POOL: Optional[ThreadPoolExecutor] = None
_DECODE_LOCK = threading.Lock()


def _decode_pool() -> ThreadPoolExecutor:
    """Process-wide JPEG decode pool: a few threads are enough to outrun the tracker (≈2 ms per 480p frame and thread)."""
    global _DECODE_POOL
    with _DECODE_LOCK:
        if _DECODE_POOL is None:
            n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
            _DECODE_POOL = ThreadPoolExecutor(max_workers=max(2, min(8, n // 2)), thread_name_prefix="frtm-decode")
        return _DECODE_POOL


def transpose_dict(d: dict) -> Dict[object, list]:
    """{key: value} -> {value: [keys...]} in insertion order (``lib/datasets.py:9-13``)."""
    out: Dict[object, list] = {}
    for k, v in d.items():
        out.setdefault(v, []).append(k)
    return out


class _Preload:
    """One in-flight preload of a sequence: decode futures, the pinned slab, the copy stream."""

    def __init__(self, seq: "FileSequence", device):
        self.device = torch.device(device)
        self.frames: List[Optional[torch.Tensor]] = [None] * len(seq.images)
        self.slab: Optional[torch.Tensor] = None
        self.slab_shape = None
        self.lock = threading.Lock()
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        pool = _decode_pool()
        self.futures = [pool.submit(self._load, seq.images[i], i, len(seq.images)) for i in range(len(seq.images))]

    def _staging(self, shape, n) -> Optional[torch.Tensor]:
        """The pinned slab is sized from the first decoded frame; frames of another size (never in DAVIS / YouTubeVOS) fall
        back to their own pinned tensor."""
        with self.lock:
            if self.slab is None:
                self.slab = torch.empty((n,) + tuple(shape), dtype=torch.uint8, pin_memory=True)
                self.slab_shape = tuple(shape)
            return self.slab if tuple(shape) == self.slab_shape else None

    def _load(self, path, i, n):
        im = imread(path)
        if not self.cuda:
            self.frames[i] = im.to(self.device)
            return
        slab = self._staging(im.shape, n)
        host = slab[i] if slab is not None else torch.empty(im.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(im)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.frames[i] = host.to(self.device, non_blocking=True)

    def wait(self):
        for f in self.futures:
            f.result()                      # re-raises decode errors on the caller's thread
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self.stream)    # device-side ordering only: the host does not block on the copies
            for t in self.frames:
                t.record_stream(cur)
        return self.frames


class FileSequence(torch.utils.data.Dataset):
    """Inference-only sequence backed by JPEG images and start-label PNGs (``lib/datasets.py:16-69``)."""

    def __init__(self, dset_name, seq_name, jpeg_path: Path, anno_path: Path, start_frames: dict, merge_objects=False,
                 all_annotations=False):
        self.dset_name = dset_name
        self.name = seq_name
        self.images = sorted(Path(jpeg_path).glob("*.jpg"))
        self.preloaded_images = None
        self.anno_path = Path(anno_path)
        self.start_frames = transpose_dict(start_frames)          # frame name -> object ids starting there
        self.obj_ids = list(start_frames.keys()) if not merge_objects else [1]
        self.frame_names = [f.stem for f in self.images]
        self.merge_objects = merge_objects
        if all_annotations:
            self.annos = sorted(self.anno_path.glob("*.png"))
        self._pending: Optional[_Preload] = None

    def __len__(self):
        return len(self.images)

    def __getitem__(self, item):
        im = self.preloaded_images[item] if self.preloaded_images is not None else imread(self.images[item])
        lb = []
        name = self.frame_name(item)
        obj_ids = self.start_frames.get(name, [])
        if len(obj_ids) > 0:
            lb = imread(self.anno_path / (name + ".png"))
            if self.merge_objects:
                lb = (lb != 0).byte()
                obj_ids = [1]
            else:
                # labels of objects that do not start on this frame are suppressed (YouTubeVOS, ``:52-56``): one pass
                # with a 256-entry table instead of one masked write per foreign id
                keep = torch.zeros(256, dtype=torch.uint8)
                keep[torch.tensor(obj_ids, dtype=torch.long)] = torch.tensor(obj_ids, dtype=torch.uint8)
                lb = keep[lb.long()]
        return im, lb, obj_ids

    def frame_name(self, item):
        return self.images[item].stem

    def preload_async(self, device):
        """Start decoding + uploading every frame in the background; ``preload`` (or the next call) collects it."""
        if self._pending is None and self.preloaded_images is None:
            self._pending = _Preload(self, device)
        return self

    def preload(self, device):
        """Preload all images and upload them to ``device`` (``:63-65``); returns when every frame is decoded and its
        copy is enqueued and ordered before the caller's current stream."""
        if self.preloaded_images is not None:
            return
        self.preload_async(device)
        self.preloaded_images = self._pending.wait()
        self._keepalive = self._pending          # the pinned slab must outlive the asynchronous copies
        self._pending = None

    def release(self):
        """Drop the preloaded frames (device memory and the pinned slab)."""
        self.preloaded_images = None
        self._pending = None
        self._keepalive = None

    def __repr__(self):
        return "%s: %s, %d frames" % (self.dset_name, self.name, len(self.images))


class _FileDataset:
    """What DAVIS and YouTubeVOS share: a root that must exist (the reference prints and calls ``quit(1)`` otherwise,
    ``:78-80`` — same message, same exit status), a sorted sequence list narrowed by ``sequences`` / ``restart``, a start
    frame per object and sequence, and ``FileSequence`` objects created on access.  The most recent one is cached so that
    the prefetch ``Tracker.run_dataset`` starts on sequence k+1 is found again when the loop gets there."""

    merge_objects = False

    def __init__(self, path, name: str, year: str, all_annotations: bool):
        self.dset_path = Path(path).expanduser().resolve()
        if not self.dset_path.exists():
            print("Dataset directory '%s' not found." % path)
            raise SystemExit(1)
        self.name, self.year, self.all_annotations = name, year, all_annotations
        self._cache = (None, None)

    def _narrow(self, sequences, restart):
        if sequences is not None:
            assert set(sequences).issubset(self.sequences)
            self.sequences = sorted(set(self.sequences).intersection(sequences))
        if restart is not None:
            assert restart in self.sequences
            self.sequences = self.sequences[self.sequences.index(restart):]

    def __len__(self):
        return len(self.sequences)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __getitem__(self, item):
        seq = self.sequences[item]                   # IndexError ends old-style iteration, as with the reference classes
        key = (seq, self.all_annotations)
        if self._cache[0] != key:
            self._cache = (key, FileSequence(self.name, seq, self.jpeg_path / seq, self.anno_path / seq, self.start_frames[seq],
                                             merge_objects=self.merge_objects, all_annotations=self.all_annotations))
        return self._cache[1]


def _read_lines(path) -> List[str]:
    with open(path) as f:
        return [line.strip() for line in sorted(f.readlines())]


class DAVISDataset(_FileDataset):
    """DAVIS 2016 / 2017 at 480p (``lib/datasets.py:72-112``): ``ImageSets/<year>/<split>.txt`` lists the sequences, every
    object starts on frame ``00000`` (its ids are read from that annotation), 2016 merges all objects into one."""

    FIRST_FRAME = "00000"

    def __init__(self, path, year: str, split: str, restart: str = None, sequences=None, all_annotations=False):
        super().__init__(path, "dv%s%s" % (year, split), year, all_annotations)
        self.jpeg_path = self.dset_path / "JPEGImages" / "480p"
        self.anno_path = self.dset_path / "Annotations" / "480p"
        self.merge_objects = year == "2016"
        self.sequences = _read_lines(self.dset_path / "ImageSets" / year / (split + ".txt"))
        self._narrow(sequences, restart)
        self.start_frames = {}
        for seq in self.sequences:
            ids = np.unique(imread(self.anno_path / seq / (self.FIRST_FRAME + ".png")).numpy())
            self.start_frames[seq] = {int(i): self.FIRST_FRAME for i in ids if i != 0}


class YouTubeVOSDataset(_FileDataset):
    """YouTubeVOS 2018 (``lib/datasets.py:115-158``): an object starts on the first frame ``meta.json`` lists for it.

    Splits: ``valid`` / ``test`` (sequences = annotation directories) and ``train`` / ``jjval`` (sequences from a list file:
    ``ytvos_jjtrain.txt`` / ``ytvos_jjvalid.txt`` ship with the reference next to its ``lib/datasets.py`` — pass the location
    as ``imset`` or place the file next to this module); the ``*_all_frames`` variants read the densely sampled JPEGs but
    the same annotations and metadata."""

    #           split  -> (annotation split, sequence list file or None)
    _SPLITS = {"train": ("train", "ytvos_jjtrain.txt"), "jjval": ("train", "ytvos_jjvalid.txt"),
               "valid": ("valid", None), "test": ("test", None)}

    def __init__(self, path, year: str, split: str, restart: str = None, sequences=None, all_annotations=False, imset=None):
        super().__init__(path, "ytvos%s%s" % (year, split), year, all_annotations)
        dense = split.endswith("_all_frames")
        base = split[:-len("_all_frames")] if dense else split
        if base not in self._SPLITS:
            raise ValueError("YouTubeVOSDataset: unknown split %r" % (split,))
        anno_split, list_file = self._SPLITS[base]
        self.jpeg_path = self.dset_path / (anno_split + "_all_frames" if dense else anno_split) / "JPEGImages"
        self.anno_path = self.dset_path / anno_split / "Annotations"
        if list_file is not None:
            imset = Path(imset) if imset is not None else Path(__file__).parent / list_file
            if not imset.exists():
                raise FileNotFoundError("YouTubeVOSDataset(split=%r): sequence list %s not found (pass imset=...)" % (split, imset))
            self.sequences = _read_lines(imset)
        else:
            self.sequences = [d.name for d in sorted(self.anno_path.glob("*")) if d.is_dir()]
        with open(self.dset_path / anno_split / "meta.json") as f:
            self.meta = json.load(f)["videos"]
        self._narrow(sequences, restart)
        self.start_frames = {seq: {int(obj_id): info["frames"][0] for obj_id, info in self.meta[seq]["objects"].items()}
                             for seq in self.sequences}
