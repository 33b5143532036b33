"""-m gpu: the sequence / dataset drivers on the device — ``run_sequence(speedrun=True)`` (the DAVIS-2016 path of
``evaluate.py:157``), the reference-API target construction (``TargetObject.initialize``), ``track`` before any object is
live, and row f2 of SURVEY.md §8: file-backed sequences whose JPEG decode + pinned-slab upload run behind the tracking
of the previous sequence (``lib/datasets.py``, ``Tracker.run_dataset``)."""
from pathlib import Path

import numpy as np
import pytest
import torch
from PIL import Image

import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SIZE = GI.MID


def _tracker(n_frames=10, n_obj=2, seq_id=3):
    from frtm_vos_b200 import synth
    from test_gpu_model import _build
    bb = synth.backbone_state_dict("resnet18", size=SIZE)
    seg = synth.segnet_state_dict("resnet18")
    trk, fe = _build("resnet18", bb, seg, GI.disc_params(256))
    seq = synth.SyntheticSequence(num_objects=n_obj, num_frames=n_frames, size=SIZE, seq_id=seq_id)
    return trk, fe, seq


def test_speedrun_warmup_matches_plain_run():
    """``speedrun=True`` initialises on frame 0, calls ``track`` while no object is live yet (the reference skips its
    per-object loops there, ``model/tracker.py:120-124,193-227``), drops the targets and runs the sequence.  The warm-up
    consumes the global torch generator the target models draw their initial weights from (as in the reference), so the
    comparison run consumes it the same way first (a one-frame sequence = the warm-up's ``initialize``): identical labels."""
    import replay
    trk, fe, seq = _tracker(n_frames=10, n_obj=1)          # DAVIS 2016: one object
    torch.manual_seed(11)
    trk.run_sequence(replay.Head(seq, 1))
    plain, _ = trk.run_sequence(seq)
    torch.manual_seed(11)
    fast, fps = trk.run_sequence(seq, speedrun=True)
    assert len(fast) == len(plain) == len(seq) and fps > 0
    for a, b in zip(fast, plain):
        assert torch.equal(a.cpu(), b.cpu())
    trk2, _, seq2 = _tracker(n_frames=9, n_obj=2)
    out, _ = trk2.run_sequence(seq2, speedrun=True)        # multi-object warm-up takes the merge of the start masks
    assert len(out) == 9


def test_track_with_only_fresh_targets_merges_start_masks():
    from oracle import frtm_ref as R
    trk, fe, seq = _tracker(n_frames=3, n_obj=2)
    image, labels, ids = seq[0]
    trk.object_ids = seq.obj_ids
    trk.targets = dict()
    trk.current_frame = 0
    trk._lut = torch.tensor([0] + list(seq.obj_ids), dtype=torch.uint8, device=DEV)
    trk.initialize(image.to(DEV), labels.to(DEV), ids)
    masks = trk.track(image.to(DEV))
    cm = torch.zeros(3, *SIZE)
    for k, oid in enumerate(ids):
        cm[k + 1] = (labels[0] == oid).float()
    ref = R.merge_masks(cm)
    assert torch.allclose(masks.cpu(), ref, atol=1e-6)
    assert torch.equal(trk._last_labels.cpu(), labels[0])


def test_target_object_initialize_reference_api():
    """``TargetObject.initialize(ft, mask)`` with the dict the public ``ResnetFeatureExtractor.__call__`` returns."""
    from frtm_vos_b200.model.tracker import TargetObject
    trk, fe, seq = _tracker(n_frames=2, n_obj=1)
    image, labels, ids = seq[0]
    mask = (labels == 1).byte().to(DEV)
    im5, m5 = trk.augment(image.to(DEV), mask)
    ft = fe(im5.to(DEV), ["layer4"])
    assert set(ft.keys()) == {"layer4"} and ft.nhwc == {}
    torch.manual_seed(0)
    t = TargetObject(obj_id=1, index=1, disc_params=trk.disc_params, start_frame=0, start_mask=mask)
    t.initialize(ft, m5.to(DEV))
    d = t.discriminator
    assert d.memory.current_size == 5 and torch.isfinite(d.filter.weight).all()
    s = d(ft["layer4"][:1])
    up = torch.nn.functional.interpolate(s, SIZE, mode="bilinear", align_corners=False)[0, 0]
    inter = ((up > 0.5) & (mask[0] > 0)).sum().item()
    union = ((up > 0.5) | (mask[0] > 0)).sum().item()
    assert inter / max(union, 1) > 0.6                       # the fitted target model segments its own first frame


# ---- row f2: file-backed dataset on the device ---------------------------------------------------------------------------
def _write_davis(root: Path, seqs):
    from frtm_vos_b200.lib.image import imwrite_indexed
    for name, seq in seqs.items():
        jd, ad = root / "JPEGImages" / "480p" / name, root / "Annotations" / "480p" / name
        jd.mkdir(parents=True); ad.mkdir(parents=True)
        for t in range(len(seq)):
            im, lb, ids = seq[t]
            Image.fromarray(im.permute(1, 2, 0).numpy()).save(jd / ("%05d.jpg" % t), quality=95)
            if t == 0:
                imwrite_indexed(ad / "00000.png", lb[0])
    (root / "ImageSets" / "2017").mkdir(parents=True)
    (root / "ImageSets" / "2017" / "val.txt").write_text("\n".join(seqs) + "\n")


class _Preloaded:
    """Plain in-memory sequence over the same decoded frames (what the reference's synchronous preload produces)."""

    def __init__(self, fs):
        self.name, self.obj_ids, self.frame_names = fs.name, fs.obj_ids, fs.frame_names
        self.items = [fs[i] for i in range(len(fs))]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        im, lb, ids = self.items[i]
        return im.to(DEV), (lb.to(DEV) if torch.is_tensor(lb) else lb), ids


def test_run_dataset_async_preload_equals_preloaded_path(tmp_path):
    from frtm_vos_b200 import synth
    from frtm_vos_b200.lib import datasets as DS
    from frtm_vos_b200.lib.image import imread
    seqs = {"s%d" % k: synth.SyntheticSequence(num_objects=2, num_frames=9 + k, size=SIZE, seq_id=20 + k) for k in range(3)}
    root = tmp_path / "DAVIS"
    _write_davis(root, seqs)
    ds = DS.DAVISDataset(root, "2017", "val")
    assert len(ds) == 3
    trk, fe, _ = _tracker(n_frames=2)
    # reference path: frames decoded on the host and handed over as tensors, one sequence at a time
    want = {}
    for k in range(len(ds)):
        plain = _Preloaded(DS.DAVISDataset(root, "2017", "val")[k])
        torch.manual_seed(11)
        out, _ = trk.run_sequence(plain)
        want[plain.name] = torch.stack([o.reshape(SIZE).cpu() for o in out])
    # driver path: threaded decode into a pinned slab, asynchronous H2D on a copy stream, next sequence prefetched
    started = []
    orig = DS.FileSequence.preload_async

    def spy(self, device):
        started.append((self.name, torch.device(device).type))
        return orig(self, device)

    DS.FileSequence.preload_async = spy
    seeds = iter([11, 11, 11])
    orig_run = trk.run_sequence

    def seeded(sequence, speedrun=False, **kw):
        torch.manual_seed(next(seeds))
        assert sequence.preloaded_images is not None and all(t.is_cuda for t in sequence.preloaded_images)
        return orig_run(sequence, speedrun, **kw)

    trk.run_sequence = seeded
    try:
        trk.run_dataset(ds, tmp_path / "out")
    finally:
        DS.FileSequence.preload_async = orig
    assert [n for n, _ in started].count("s1") >= 1 and all(t == "cuda" for _, t in started)
    for name, ref in want.items():
        files = sorted((tmp_path / "out" / name).glob("*.png"))
        assert len(files) == ref.shape[0]
        got = torch.stack([imread(f)[0] for f in files])
        assert torch.equal(got, ref), name


def test_block_graph_is_captured_once_and_matches_eager():
    """``graph_blocks``: blocks aligned to the update schedule are replayed as one CUDA graph per block length (full 8-frame
    blocks and the shorter tail of a sequence).  With the per-slot target-model buffers pooled across sequences the graphs
    captured on the first sequence serve the following ones (same object count): identical labels, filters and memory
    weights to the eager path, one capture per block length for three sequences."""
    from frtm_vos_b200 import synth
    trk_e, _, _ = _tracker()
    trk_e.graph_blocks = False
    trk_g, _, _ = _tracker()
    trk_g.graph_blocks = True
    seqs = [synth.SyntheticSequence(num_objects=2, num_frames=29, size=SIZE, seq_id=30 + k) for k in range(3)]
    for seq in seqs:                                     # 28 tracked frames: three full blocks + a 4-frame tail
        torch.manual_seed(11)
        a, _ = trk_e.run_sequence(seq)
        torch.manual_seed(11)
        b, _ = trk_g.run_sequence(seq)
        for x, y in zip(a, b):
            assert torch.equal(x.cpu(), y.cpu()), seq.name
        for o in seq.obj_ids:
            de, dg = trk_e.targets[o].discriminator, trk_g.targets[o].discriminator
            assert torch.equal(de.filter.weight.cpu(), dg.filter.weight.cpu())
            assert torch.equal(de.memory.weights.cpu(), dg.memory.weights.cpu())
    assert trk_g.graph_captures == 2, trk_g.graph_captures
    assert trk_e.graph_captures == 0


def test_prefetched_initialisation_gives_identical_sequences():
    """``run_sequence(seq, next_sequence=nxt)`` prepares the first-frame augmentations of ``nxt`` in worker threads behind
    the tracking of ``seq`` (``Tracker.prefetch_init``); ``nxt`` must then run exactly as without the overlap: same views,
    hence bit-identical labels, filters and memory weights.  Device-resident and host (pinned) sequences."""
    from frtm_vos_b200 import synth
    trk, fe, seq_a = _tracker(n_frames=18, n_obj=2, seq_id=3)
    seq_b = synth.SyntheticSequence(num_objects=3, num_frames=18, size=SIZE, seq_id=5)

    def run(seq, nxt):
        torch.manual_seed(5)
        outs, _ = trk.run_sequence(seq, next_sequence=nxt)
        filt = [trk.targets[o].discriminator.filter.weight.detach().clone() for o in seq.obj_ids]
        wts = [trk.targets[o].discriminator.memory.weights.clone() for o in seq.obj_ids]
        return [o.cpu() for o in outs], filt, wts

    trk.prefetch_next = False
    ref_b = run(seq_b, None)
    trk.prefetch_next = True
    run(seq_a, seq_b)                                     # prepares seq_b's three objects
    assert (id(seq_b), 0) in trk._prefetched and len(trk._prefetched[(id(seq_b), 0)]["futures"]) == 3
    got_b = run(seq_b, seq_a)                             # consumes them (and prepares seq_a's)
    assert (id(seq_b), 0) not in trk._prefetched
    for a, b in zip(ref_b[0], got_b[0]):
        assert torch.equal(a, b)
    for a, b in zip(ref_b[1] + ref_b[2], got_b[1] + got_b[2]):
        assert torch.equal(a, b)

    class Host:                                           # frames in pinned host memory: the prefetch uploads frame 0 itself
        def __init__(self, seq):
            self.name, self.obj_ids, self.frame_names = seq.name, seq.obj_ids, seq.frame_names
            self.items = [(im.pin_memory(), (lb.pin_memory() if torch.is_tensor(lb) else lb), ids)
                          for im, lb, ids in (seq[t] for t in range(len(seq)))]

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return self.items[i]

    hb = Host(seq_b)
    run(seq_a, hb)
    got_h = run(hb, None)
    for a, b in zip(ref_b[0], got_h[0]):
        assert torch.equal(a, b)
    for a, b in zip(ref_b[1] + ref_b[2], got_h[1] + got_h[2]):
        assert torch.equal(a, b)
    # label maps streamed to pinned host memory block by block
    pinned = torch.zeros((len(seq_b), *SIZE), dtype=torch.uint8).pin_memory()
    torch.manual_seed(5)
    outs, _ = trk.run_sequence(seq_b, host_labels=pinned)
    assert torch.equal(torch.stack([o.reshape(SIZE).cpu() for o in outs]), pinned)
    assert torch.equal(pinned, torch.stack([o.reshape(SIZE) for o in ref_b[0]]))
    # chained sequences: no device synchronisation between them, the joint fits of seq_b's objects start while seq_a's
    # last block is still running; complete once ``sequence_done`` has been reached
    pinned.zero_()
    torch.manual_seed(5)
    trk.run_sequence(seq_a, next_sequence=seq_b, sync=False)
    torch.manual_seed(5)
    outs, fps = trk.run_sequence(seq_b, host_labels=pinned, sync=False)
    assert fps != fps                                      # nan: not measured
    trk.sequence_done.synchronize()
    assert torch.equal(pinned, torch.stack([o.reshape(SIZE) for o in ref_b[0]]))
    got = [trk.targets[o].discriminator.filter.weight.detach().clone() for o in seq_b.obj_ids] + \
          [trk.targets[o].discriminator.memory.weights.clone() for o in seq_b.obj_ids]
    for a, b in zip(ref_b[1] + ref_b[2], got):
        assert torch.equal(a, b)


def test_block_inserts_may_skip_the_fullres_mirrors():
    """``Tracker.store_fullres_memory``: with it the block inserts keep ``Memory.labels`` / ``Memory.pixel_weights`` up to date
    (reference buffers, model/memory.py:20-21); without it (default) those two copies are skipped.  Nothing on the path reads
    them, so labels, filters, samples, stencils and sample weights are identical either way."""
    from frtm_vos_b200 import synth
    seq = synth.SyntheticSequence(num_objects=2, num_frames=18, size=SIZE, seq_id=51)
    res = {}
    for full in (False, True):
        trk, _, _ = _tracker(n_frames=2)
        trk.store_fullres_memory = full
        torch.manual_seed(13)
        outs, _ = trk.run_sequence(seq)
        mems = [trk.targets[o].discriminator.memory for o in seq.obj_ids]
        res[full] = ([o.cpu() for o in outs], [trk.targets[o].discriminator.filter.weight.detach().clone() for o in seq.obj_ids],
                     [(m.samples.clone(), m.stencil.clone(), m.uty.clone(), m.weights.clone(), m.split.clone()) for m in mems],
                     [(m.labels.clone(), m.pixel_weights.clone(), int(m.state[0])) for m in mems])
    for a, b in zip(res[False][0], res[True][0]):
        assert torch.equal(a, b)
    for a, b in zip(res[False][1], res[True][1]):
        assert torch.equal(a, b)
    for ma, mb in zip(res[False][2], res[True][2]):
        for a, b in zip(ma, mb):
            assert torch.equal(a, b)
    for (la, pa, na), (lb, pb, nb) in zip(res[False][3], res[True][3]):
        assert na == nb and na > 5                                  # the block inserts happened
        assert torch.equal(la[:5], lb[:5]) and torch.equal(pa[:5], pb[:5])      # first-frame samples: always stored
        filled = lambda x: int((x[5:na].flatten(1).abs().sum(1) > 0).sum())
        assert filled(lb) == na - 5 and filled(pb) == na - 5        # every inserted sample mirrored
        assert filled(la) <= 1 and filled(pa) <= 1                  # block inserts skipped (the 17th frame runs frame by frame)

