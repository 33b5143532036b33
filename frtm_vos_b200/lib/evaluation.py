"""Dataset-level J / F evaluation of written label maps (``lib/evaluation.py:9-85``): same
console / ``evaluation-<measure>.txt`` report, line for line, as the reference."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import davis as utils
from .image import imread

_BLOCKS = np.array(("u", " ", "▁", "▂", "▃", "▄", "▅", "▆", "▇", "█", "o"))


def text_bargraph(values) -> str:
    """One character per value: eighth-block bars for [0, 1], 'u' below, 'o' above, '░' for NaN (``lib/utils.py:9-22``)."""
    v = np.array(values, dtype=np.float64)
    nans = np.isnan(v)
    v[nans] = 0
    nsteps = len(_BLOCKS) - 3
    idx = ((v + 1 / (2 * nsteps)) * nsteps + 1).astype(np.int64)
    idx[v < 0] = 0
    idx[v > 1] = len(_BLOCKS) - 1
    graph = _BLOCKS[idx]
    graph[nans] = "░"
    return "".join(graph)


def evaluate_dataset(dset, results_path, measure="J", to_file=True):
    """Scores every sequence of ``dset`` (built with ``all_annotations=True``) against the PNGs under ``results_path`` and
    prints / writes the report.  Returns ``{sequence name: evaluate_sequence result}`` (the reference returns nothing)."""
    results = OrderedDict()
    dset_scores, dset_decay, dset_recall = [], [], []
    f = open(results_path / ("evaluation-%s.txt" % measure), "w") if to_file else None

    def _print(msg):
        print(msg)
        if f is not None:
            print(msg, file=f)
            f.flush()

    try:
        n_seqs = len(dset)
        for j, sequence in enumerate(dset):
            annotations, segmentations = OrderedDict(), OrderedDict()
            for file in sequence.annos:
                lb = imread(file)
                annotations[file.stem] = (lb != 0).byte() if sequence.merge_objects else lb
                segmentations[file.stem] = imread(results_path / sequence.name / file.name)

            object_info = dict()
            for obj_id in sequence.obj_ids:
                for frame, obj_ids in sequence.start_frames.items():
                    if obj_id in obj_ids:
                        assert obj_id not in object_info          # one start frame per object
                        object_info[obj_id] = frame
            assert 0 not in object_info

            n_objs = len(object_info)
            _print("%d/%d: %s: %d object%s" % (j + 1, n_seqs, sequence.name, n_objs, "s" if n_objs > 1 else ""))
            r = utils.evaluate_sequence(segmentations, annotations, object_info, measure=measure)
            results[sequence.name] = r

            per_obj_score, per_frame_score = [], []
            for obj_id, score in r["raw"].items():
                per_frame_score.append(score)
                s = utils.mean(score)
                per_obj_score.append(s)
                if n_objs > 1:
                    _print("joint {obj}: acc {score:.3f} ┊{apf}┊".format(obj=obj_id, score=s, apf=text_bargraph(score)))

            dset_decay.extend(r["decay"])
            dset_recall.extend(r["recall"])
            dset_scores.extend(per_obj_score)
            seq_score = utils.mean(per_obj_score)
            seq_mean_score = utils.nanmean(np.array(per_frame_score), axis=0)
            _print("final  : acc {seq:.3f} ({dset:.3f}) ┊{apf}┊".format(seq=seq_score, dset=np.mean(dset_scores),
                                                                         apf=text_bargraph(seq_mean_score)))
        _print("%s: %.3f, recall: %.3f, decay: %.3f" % (measure, utils.mean(dset_scores), utils.mean(dset_recall),
                                                        utils.mean(dset_decay)))
    finally:
        if f is not None:
            f.close()
    return results
