"""Dataset-level J / F evaluation of written label maps.

``evaluate_dataset(dset, results_path, measure, to_file)`` keeps the reference's call and its report — console and
``evaluation-<measure>.txt``, line for line (``lib/evaluation.py:9-85``) — but is organised as two stages: sequences are
*scored* independently (label maps read once per sequence, all objects of a sequence measured from the same arrays; with
``workers > 1`` several sequences at a time on a thread pool — OpenCV and numpy release the GIL), and a *reporter* consumes the
scores strictly in dataset order, so the running dataset average printed after every sequence does not depend on the
scheduling.  The scores are also returned (the reference returns nothing).
"""
from __future__ import annotations

from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, Iterator, List

import numpy as np

from . import davis
from .image import imread

_BAR_GLYPHS = np.array(("u", " ", "▁", "▂", "▃", "▄", "▅", "▆", "▇", "█", "o"))
_BAR_LEVELS = len(_BAR_GLYPHS) - 3           # glyphs 1..9 cover [0, 1] in eighths; 0 / 10 flag values outside


def text_bargraph(values) -> str:
    """One glyph per value: eighth-block bars for [0, 1], 'u' below 0, 'o' above 1, '░' for NaN (``lib/utils.py:9-22``)."""
    v = np.array(values, dtype=np.float64)
    missing = np.isnan(v)
    v[missing] = 0.0
    level = ((v + 0.5 / _BAR_LEVELS) * _BAR_LEVELS + 1).astype(np.int64)
    level[v < 0] = 0
    level[v > 1] = len(_BAR_GLYPHS) - 1
    glyphs = _BAR_GLYPHS[level]
    glyphs[missing] = "░"
    return "".join(glyphs)


@dataclass
class SequenceScore:
    """Scores of one sequence: ``result`` is ``davis.evaluate_sequence``'s dict, ``per_object`` the NaN-ignoring mean of each
    object's per-frame scores (object order = ``sequence.obj_ids`` order)."""
    name: str
    result: dict
    per_object: List[float] = field(default_factory=list)

    @property
    def n_objects(self) -> int:
        return len(self.result["raw"])

    @property
    def mean(self) -> float:
        return davis.mean(self.per_object)

    @property
    def per_frame(self) -> np.ndarray:
        """Mean over the objects for every frame (NaN where no object is scored)."""
        return davis.nanmean(np.array(list(self.result["raw"].values())), axis=0)


def _start_frame_of_each_object(sequence) -> Dict[int, str]:
    starts: Dict[int, str] = {}
    for frame, ids in sequence.start_frames.items():
        for obj_id in ids:
            if obj_id in sequence.obj_ids:
                assert obj_id not in starts, "object %r starts on more than one frame" % (obj_id,)
                starts[obj_id] = frame
    assert 0 not in starts, "label 0 is the background"
    return {obj_id: starts[obj_id] for obj_id in sequence.obj_ids if obj_id in starts}


def score_sequence(sequence, results_path, measure: str) -> SequenceScore:
    """Reads every annotated frame of ``sequence`` and the label map written for it under ``results_path/<name>/`` and
    measures each object from its start frame on."""
    truth, written = OrderedDict(), OrderedDict()
    for path in sequence.annos:
        lb = imread(path)
        truth[path.stem] = (lb != 0).byte() if sequence.merge_objects else lb
        written[path.stem] = imread(results_path / sequence.name / path.name)
    result = davis.evaluate_sequence(written, truth, _start_frame_of_each_object(sequence), measure=measure)
    return SequenceScore(sequence.name, result, [davis.mean(s) for s in result["raw"].values()])


def iter_scores(dset, results_path, measure: str, workers: int = 1) -> Iterator[SequenceScore]:
    """Scores in dataset order; with ``workers > 1`` up to that many sequences are scored concurrently."""
    if workers <= 1:
        for sequence in dset:
            yield score_sequence(sequence, results_path, measure)
        return
    with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="frtm-eval") as pool:
        pending = []
        for sequence in dset:
            pending.append(pool.submit(score_sequence, sequence, results_path, measure))
            if len(pending) >= 2 * workers:
                yield pending.pop(0).result()
        for fut in pending:
            yield fut.result()


class _Reporter:
    """Formats the per-sequence lines and keeps the dataset-level accumulators (one entry per object)."""

    def __init__(self, measure: str, n_sequences: int, emit):
        self.measure, self.n, self.emit = measure, n_sequences, emit
        self.seen = 0
        self.scores: List[float] = []
        self.recall: List[float] = []
        self.decay: List[float] = []

    def sequence(self, s: SequenceScore):
        self.seen += 1
        self.emit("%d/%d: %s: %d object%s" % (self.seen, self.n, s.name, s.n_objects, "" if s.n_objects == 1 else "s"))
        if s.n_objects > 1:
            for (obj_id, frames), acc in zip(s.result["raw"].items(), s.per_object):
                self.emit("joint %s: acc %.3f ┊%s┊" % (obj_id, acc, text_bargraph(frames)))
        self.scores += s.per_object
        self.recall += s.result["recall"]
        self.decay += s.result["decay"]
        self.emit("final  : acc %.3f (%.3f) ┊%s┊" % (s.mean, np.mean(self.scores), text_bargraph(s.per_frame)))

    def summary(self):
        self.emit("%s: %.3f, recall: %.3f, decay: %.3f" % (self.measure, davis.mean(self.scores), davis.mean(self.recall),
                                                           davis.mean(self.decay)))


def evaluate_dataset(dset, results_path, measure="J", to_file=True, workers: int = 1):
    """Scores every sequence of ``dset`` (built with ``all_annotations=True``) against the PNGs under ``results_path``,
    prints the report and, with ``to_file``, writes it to ``results_path/evaluation-<measure>.txt``.
    Returns ``{sequence name: davis.evaluate_sequence result}``."""
    sink = open(results_path / ("evaluation-%s.txt" % measure), "w") if to_file else None

    def emit(line):
        print(line)
        if sink is not None:
            print(line, file=sink, flush=True)

    out = OrderedDict()
    try:
        report = _Reporter(measure, len(dset), emit)
        for score in iter_scores(dset, results_path, measure, workers):
            report.sequence(score)
            out[score.name] = score.result
        report.summary()
    finally:
        if sink is not None:
            sink.close()
    return out
