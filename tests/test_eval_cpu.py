"""CPU tests of the rows SURVEY.md §8(f) marks "next": f2 (file-backed datasets + the run_dataset driver) and f3 (J / F
evaluation).  Parity is against the executed reference where its tree is present (``needs_reference``) and against the
fixtures it produced (``tests/golden/eval.npz``, ``oracle/make_golden.py eval``) everywhere else."""
import json
import warnings
from pathlib import Path
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch
from PIL import Image

import golden_inputs as GI
from frtm_vos_b200.lib import davis as D
from frtm_vos_b200.lib import datasets as DS
from frtm_vos_b200.lib.evaluation import evaluate_dataset, text_bargraph
from frtm_vos_b200.lib.image import imread, imwrite_indexed

H, W = 48, 64


def _blob(rng, h=H, w=W):
    yy, xx = np.mgrid[:h, :w]
    cy, cx = rng.uniform(0.25, 0.75) * h, rng.uniform(0.25, 0.75) * w
    return ((yy - cy) / (0.18 * h)) ** 2 + ((xx - cx) / (0.15 * w)) ** 2 <= 1


def _write_sequence(jpeg_dir: Path, anno_dir: Path, n_frames, obj_ids, rng, annotate="all", start=None):
    """Frames 00000.. as JPEG; palette PNG annotations (all frames or only the frames in ``start``: {obj_id: frame index})."""
    jpeg_dir.mkdir(parents=True)
    anno_dir.mkdir(parents=True)
    blobs = {o: _blob(rng) for o in obj_ids}
    for t in range(n_frames):
        Image.fromarray(rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)).save(jpeg_dir / ("%05d.jpg" % t), quality=90)
        lb = np.zeros((H, W), np.uint8)
        for o in obj_ids:
            if start is None or t >= start[o]:
                lb[np.roll(blobs[o], t, 1)] = o
        if annotate == "all" or (start is not None and t in start.values()):
            imwrite_indexed(anno_dir / ("%05d.png" % t), torch.from_numpy(lb))


@pytest.fixture(scope="module")
def davis_tree(tmp_path_factory):
    root = tmp_path_factory.mktemp("DAVIS")
    rng = np.random.RandomState(1)
    seqs = dict(alpha=(6, [1, 2]), beta=(5, [1, 2, 3]), gamma=(7, [1]))
    for name, (n, ids) in seqs.items():
        _write_sequence(root / "JPEGImages" / "480p" / name, root / "Annotations" / "480p" / name, n, ids, rng)
    for year in ("2016", "2017"):
        (root / "ImageSets" / year).mkdir(parents=True)
        (root / "ImageSets" / year / "val.txt").write_text("gamma\nalpha\nbeta\n")
    return root


@pytest.fixture(scope="module")
def ytvos_tree(tmp_path_factory):
    root = tmp_path_factory.mktemp("YTVOS")
    rng = np.random.RandomState(2)
    meta = dict(videos={})
    for name, n, start in (("aa11", 6, {1: 0, 2: 2}), ("bb22", 5, {3: 0}), ("cc33", 7, {1: 1, 2: 1, 5: 4})):
        _write_sequence(root / "valid" / "JPEGImages" / name, root / "valid" / "Annotations" / name, n, list(start), rng,
                        annotate="start", start=start)
        meta["videos"][name] = dict(objects={str(o): dict(frames=["%05d" % t for t in range(s, n)]) for o, s in start.items()})
    (root / "valid" / "meta.json").write_text(json.dumps(meta))
    return root


def _same_item(a, b):
    (im_a, lb_a, ids_a), (im_b, lb_b, ids_b) = a, b
    assert torch.equal(im_a, im_b) and im_a.dtype == torch.uint8 and im_a.dim() == 3
    assert list(ids_a) == list(ids_b)
    if isinstance(lb_b, list):
        assert isinstance(lb_a, list) and lb_a == lb_b
    else:
        assert lb_a.dtype == lb_b.dtype and torch.equal(lb_a, lb_b)


def _same_dataset(mine, ref):
    assert mine.name == ref.name and mine.sequences == ref.sequences and len(mine) == len(ref)
    assert mine.start_frames == ref.start_frames
    for k in range(len(ref)):
        a, b = mine[k], ref[k]
        assert (a.name, a.dset_name, a.obj_ids, a.frame_names, a.merge_objects) == (b.name, b.dset_name, b.obj_ids, b.frame_names, b.merge_objects)
        assert a.start_frames == b.start_frames and len(a) == len(b) and repr(a) == repr(b)
        assert [p.name for p in a.images] == [p.name for p in b.images]
        if hasattr(b, "annos"):
            assert [p.name for p in a.annos] == [p.name for p in b.annos]
        for i in range(len(b)):
            _same_item(a[i], b[i])
        a.preload("cpu"); b.preload("cpu")
        for i in range(len(b)):
            _same_item(a[i], b[i])


@pytest.mark.needs_reference
@pytest.mark.parametrize("year", ["2016", "2017"])
def test_davis_dataset_matches_reference(davis_tree, year):
    from oracle import shims
    ref = shims.load_reference()
    for kw in (dict(), dict(all_annotations=True), dict(sequences=["beta", "alpha"]), dict(restart="beta")):
        _same_dataset(DS.DAVISDataset(davis_tree, year, "val", **kw), ref.datasets.DAVISDataset(davis_tree, year, "val", **kw))


@pytest.mark.needs_reference
def test_ytvos_dataset_matches_reference(ytvos_tree):
    from oracle import shims
    ref = shims.load_reference()
    for split in ("valid", "valid_all_frames"):
        if split == "valid_all_frames":
            (ytvos_tree / split).mkdir(exist_ok=True)
            if not (ytvos_tree / split / "JPEGImages").exists():
                (ytvos_tree / split / "JPEGImages").symlink_to(ytvos_tree / "valid" / "JPEGImages")
        _same_dataset(DS.YouTubeVOSDataset(ytvos_tree, "2018", split), ref.datasets.YouTubeVOSDataset(ytvos_tree, "2018", split))
        _same_dataset(DS.YouTubeVOSDataset(ytvos_tree, "2018", split, restart="bb22", all_annotations=True),
                      ref.datasets.YouTubeVOSDataset(ytvos_tree, "2018", split, restart="bb22", all_annotations=True))


def test_ytvos_start_frames_and_label_suppression(ytvos_tree):
    """Objects start on the first frame listed in meta.json; on a start frame the labels of objects that do NOT start there
    are zeroed (lib/datasets.py:52-56)."""
    ds = DS.YouTubeVOSDataset(ytvos_tree, "2018", "valid")
    assert ds.sequences == ["aa11", "bb22", "cc33"] and ds.name == "ytvos2018valid"
    seq = ds[2]
    assert seq.obj_ids == [1, 2, 5] and seq.start_frames == {"00001": [1, 2], "00004": [5]}
    assert seq[0][1] == [] and seq[0][2] == []
    _, lb1, ids1 = seq[1]
    assert ids1 == [1, 2] and set(lb1.unique().tolist()) <= {0, 1, 2}
    _, lb4, ids4 = seq[4]
    raw = imread(ytvos_tree / "valid" / "Annotations" / "cc33" / "00004.png")
    assert ids4 == [5] and set(raw.unique().tolist()) == {0, 1, 2, 5}            # the file holds all three objects ...
    assert torch.equal(lb4, torch.where(raw == 5, raw, torch.zeros_like(raw)))    # ... the item only the starting one
    with pytest.raises(ValueError):
        DS.YouTubeVOSDataset(ytvos_tree, "2018", "nonsense")
    with pytest.raises(FileNotFoundError, match="imset"):
        (ytvos_tree / "train").mkdir(exist_ok=True)
        DS.YouTubeVOSDataset(ytvos_tree, "2018", "jjval")


def test_missing_dataset_directory_exits_like_the_reference(tmp_path, capsys):
    with pytest.raises(SystemExit) as e:
        DS.DAVISDataset(tmp_path / "nope", "2017", "val")
    assert e.value.code == 1 and "not found" in capsys.readouterr().out


def test_background_preload_is_ordered_and_idempotent(davis_tree):
    ds = DS.DAVISDataset(davis_tree, "2017", "val")
    seq = ds[0]
    direct = [imread(p) for p in seq.images]
    assert seq.preload_async("cpu") is seq and seq.preload_async("cpu") is seq      # second call: no second pool job set
    seq.preload("cpu")
    assert len(seq.preloaded_images) == len(direct) and all(torch.equal(a, b) for a, b in zip(seq.preloaded_images, direct))
    first = seq.preloaded_images
    seq.preload("cpu")
    assert seq.preloaded_images is first
    assert ds[0] is seq                          # the dataset hands back the prefetched object ...
    seq.release()
    assert seq.preloaded_images is None
    assert ds[1] is not seq and ds[1] is ds[1]   # ... and keeps only the most recent sequence alive
    bad = DS.FileSequence("x", "y", davis_tree / "JPEGImages" / "480p" / "alpha", davis_tree / "Annotations" / "480p" / "alpha", {1: "00000"})
    bad.images = bad.images + [davis_tree / "missing.jpg"]
    with pytest.raises(FileNotFoundError):       # decode errors surface on the caller's thread
        bad.preload("cpu")


def test_run_dataset_prefetches_the_next_sequence_and_writes_indexed_pngs(davis_tree, tmp_path):
    """Tracker.run_dataset (model/tracker.py:68-101) with a stub in place of the GPU work: sequence k+1 is being preloaded
    while sequence k runs, every frame is written as a palette PNG with the DAVIS colour map, restart skips ahead."""
    from frtm_vos_b200.model.tracker import Tracker
    from frtm_vos_b200.lib.image import davis_palette
    ds = DS.DAVISDataset(davis_tree, "2017", "val")
    events = []
    orig = DS.FileSequence.preload_async

    def spy(self, device):
        if self._pending is None and self.preloaded_images is None:
            events.append(("prefetch", self.name))
        return orig(self, device)

    def run_sequence(sequence, speedrun=False, **kw):
        events.append(("run", sequence.name))
        assert sequence.preloaded_images is not None
        outs = []
        for i in range(len(sequence)):
            lb = torch.zeros(1, H, W, dtype=torch.uint8)
            lb[0, i:i + 5, :7] = sequence.obj_ids[-1]
            outs.append(lb)
        return outs, 100.0

    stub = NS(device="cpu", clear=lambda: None, run_sequence=run_sequence)
    DS.FileSequence.preload_async = spy
    try:
        Tracker.run_dataset(stub, ds, tmp_path / "out")
    finally:
        DS.FileSequence.preload_async = orig
    assert events == [("prefetch", "alpha"), ("prefetch", "beta"), ("run", "alpha"), ("prefetch", "gamma"), ("run", "beta"),
                      ("run", "gamma")]
    for name, n, last in (("alpha", 6, 2), ("beta", 5, 3), ("gamma", 7, 1)):
        files = sorted((tmp_path / "out" / name).glob("*.png"))
        assert [f.stem for f in files] == ["%05d" % t for t in range(n)]
        im = Image.open(files[2])
        assert im.mode == "P" and np.array_equal(np.array(im.getpalette()).reshape(-1, 3)[:22], davis_palette[:22])
        assert set(np.unique(np.array(im)).tolist()) == {0, last}
    events.clear()
    Tracker.run_dataset(stub, ds, tmp_path / "out2", restart="beta")
    assert [e for e in events if e[0] == "run"] == [("run", "beta"), ("run", "gamma")]


# ---- f3: measures ---------------------------------------------------------------------------------------------------------
def test_measures_match_reference_fixtures(golden):
    g = golden("eval")
    pairs = GI.eval_mask_pairs()
    J = np.array([float(D.davis_jaccard_measure(a, b)) for a, b in pairs])
    F = np.array([float(D.davis_f_measure(a, b)) for a, b in pairs])
    assert np.array_equal(J, g["J"]) and np.array_equal(F, g["F"])              # bit-identical to the executed reference
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stats = np.array([[float(D.mean(v)), float(D.recall(v)), float(D.decay(v)), float(D.std(v))] for v in GI.eval_score_vectors()])
    assert np.array_equal(stats, g["stats"], equal_nan=True)
    bars = [text_bargraph(np.concatenate((v, [-0.1, 1.2]))) for v in GI.eval_score_vectors()]
    assert bars == [str(b) for b in g["bars"]]


def test_measure_properties():
    rng = np.random.RandomState(7)
    a = _blob(rng, 60, 80)
    z = np.zeros_like(a)
    assert D.davis_jaccard_measure(a, a) == 1 and D.davis_jaccard_measure(z, z) == 1 and D.davis_jaccard_measure(a, z) == 0
    assert D.davis_f_measure(a, a) == 1 and D.davis_f_measure(z, z) == 1 and D.davis_f_measure(a, z) == 0
    b = np.roll(a, 1, 1)                                     # a one-pixel shift is inside the boundary tolerance
    assert D.davis_f_measure(a, b) == 1 and 0.8 < D.davis_jaccard_measure(a, b) < 1
    far = np.roll(a, 30, 1)
    assert D.davis_f_measure(a, far) < 0.5
    for x, y in ((a, b), (a, far), (b, far)):                # both measures are symmetric
        assert D.davis_jaccard_measure(x, y) == D.davis_jaccard_measure(y, x)
        assert D.davis_f_measure(x, y) == pytest.approx(D.davis_f_measure(y, x), abs=1e-15)
    # integer / uint8 / torch inputs are accepted and the inputs are not modified
    a8 = a.astype(np.uint8) * 3
    keep = a8.copy()
    assert D.davis_jaccard_measure(a8, torch.from_numpy(b)) == D.davis_jaccard_measure(a, b)
    assert D.davis_f_measure(a8, b) == D.davis_f_measure(a, b) and np.array_equal(a8, keep)
    bm = D.seg2bmap(a)
    assert bm.dtype == np.bool_ and bm.shape == a.shape and not bm[-1, -1] and 0 < bm.sum() < a.sum()
    with pytest.raises(NotImplementedError):
        D.seg2bmap(a, width=40, height=30)


def test_boundary_dilation_equals_scipy_binary_dilation():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.RandomState(9)
    for (h, w) in ((37, 53), (96, 128), (12, 9)):
        m = D.seg2bmap(_blob(rng, h, w)) | (rng.uniform(size=(h, w)) > 0.995)
        m[0, 0] = m[-1, -1] = True                           # structuring element hanging over the border
        for r in (1.0, 2.0, np.ceil(0.008 * np.linalg.norm((h, w))), 5.0):
            se = D._disk(float(r))
            assert se.shape == (2 * int(r) + 1,) * 2 and se[int(r), 0] == 1 and se[0, 0] == 0 or r < 2
            assert np.array_equal(D._dilate(m, se), ndi.binary_dilation(m, structure=se))


def _results_tree(dataset, out: Path, rng):
    """Label maps a tracker could have written: the annotations, shifted and with an object dropped now and then."""
    for seq in dataset:
        (out / seq.name).mkdir(parents=True, exist_ok=True)
        for t, f in enumerate(seq.annos):
            lb = imread(f)
            if seq.merge_objects:
                lb = (lb != 0).byte()
            lb = torch.roll(lb, int(rng.randint(-2, 3)), 2)
            if t % 4 == 3:
                lb[lb == int(lb.max())] = 0
            imwrite_indexed(out / seq.name / f.name, lb)


@pytest.mark.needs_reference
@pytest.mark.parametrize("year", ["2016", "2017"])
def test_evaluate_dataset_report_matches_reference(davis_tree, tmp_path, year, capsys):
    from oracle import shims
    ref = shims.load_reference()
    mine = DS.DAVISDataset(davis_tree, year, "val", all_annotations=True)
    theirs = ref.datasets.DAVISDataset(davis_tree, year, "val", all_annotations=True)
    out = tmp_path / "results"
    _results_tree(mine, out, np.random.RandomState(4))
    for measure in ("J", "F"):
        res = evaluate_dataset(mine, out, measure=measure)
        report = (out / ("evaluation-%s.txt" % measure)).read_text()
        printed = capsys.readouterr().out
        assert printed == report and list(res) == mine.sequences
        with ref.numpy1_aliases(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.evaluation.evaluate_dataset(theirs, out, measure=measure)
        assert capsys.readouterr().out == report                                     # the reference prints what it writes
        assert (out / ("evaluation-%s.txt" % measure)).read_text() == report        # line for line, digit for digit
        assert report.splitlines()[-1].startswith("%s: " % measure) and "recall" in report
    # concurrent scoring: same report, same order
    evaluate_dataset(mine, out, measure="F", workers=3)
    assert capsys.readouterr().out == report == (out / "evaluation-F.txt").read_text()
    # per-object raw scores: NaN on the start frame and on the last frame, finite in between
    r = evaluate_dataset(mine, out, measure="J", to_file=False)[mine.sequences[0]]
    for scores in r["raw"].values():
        assert np.isnan(scores[0]) and np.isnan(scores[-1]) and np.isfinite(scores[1:-1]).all()


def test_evaluate_sequence_respects_start_frames():
    """An object that starts on frame k is scored on frames k+1 .. n-2 only (lib/davis.py:37-43)."""
    rng = np.random.RandomState(5)
    names = ["%05d" % t for t in range(8)]
    ann = {n: torch.from_numpy(np.where(_blob(rng), 2, 0).astype(np.uint8))[None] for n in names}
    seg = {n: a.clone() for n, a in ann.items()}
    seg["00005"] = torch.zeros_like(seg["00005"])
    r = D.evaluate_sequence(seg, ann, {2: "00003"}, measure="J")
    raw = r["raw"][2]
    assert np.isnan(raw[:4]).all() and np.isnan(raw[7]) and list(raw[4:7]) == [1.0, 0.0, 1.0]
    assert r["mean"] == [pytest.approx(2 / 3)] and r["recall"] == [pytest.approx(2 / 3)] and len(r["decay"]) == 1
