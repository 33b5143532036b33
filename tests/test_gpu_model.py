"""-m gpu: model-level parity (SURVEY.md §4 protocol).

P1  feed-forward at fixed state: backbone maps, projected sample, scores, logits <= 1e-3, labels identical
P4  end-to-end oracle-replay: post-init target-model state injected from the oracle, then the whole sequence
P5  end-to-end free-running: functional agreement with the oracle (init is chaotic at ulp level, finding 8)
The oracle (oracle/frtm_ref.py) runs live on the box's CPU on the same seeded inputs; the P1 case is also checked
against the committed fixture produced by the executed reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(arch, bb, seg, dp):
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200.model.seg_network import SegNetwork
    from frtm_vos_b200.model.tracker import Tracker
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    fe = ResnetFeatureExtractor(arch, state_dict=bb).to(DEV)
    chans = fe.get_out_channels()
    refiner = SegNetwork(1, 64, {L: c for L, c in chans.items() if L in ("layer5", "layer4", "layer3", "layer2")}, True)
    trk = Tracker(ImageAugmenter(GI.AUG_PARAMS), fe, dict(dp, device=DEV), refiner, DEV)
    missing = trk.load_state_dict(seg)
    trk.to(DEV)
    return trk, fe


FULL = (480, 854)      # BASELINE.json configs 2-4


@pytest.mark.parametrize("arch,size,n_obj", [("resnet18", GI.MID, 2), ("resnet101", GI.MID, 2),
                                             ("resnet18", FULL, 3), ("resnet101", FULL, 5)])
def test_P1_feedforward_fixed_state(arch, size, n_obj, golden):
    """Feed-forward at fixed (P, F): the 128x224 cases also against the executed-reference fixture, the 480x854 cases at
    the object counts of BASELINE configs 2 (rn18, 3 objects) and 3 (rn101, 5 objects) against the live oracle."""
    from oracle import frtm_ref as R
    from frtm_vos_b200 import ops
    case = GI.feedforward_case(arch, size=size, n_obj=n_obj)
    C = case["PF"][0][0].shape[1]
    trk, fe = _build(arch, case["bb"], case["seg"], GI.disc_params(C))
    img = case["image"]
    feats = fe(img.to(DEV))
    ref = R.backbone_features(case["bb"], arch, img)
    g = golden("feedforward_" + arch) if size == GI.MID else None
    for L in ("layer1", "layer2", "layer3", "layer4", "layer5"):
        d = (feats[L].cpu() - ref[L]).abs().max().item()
        assert d < 2e-4 * max(1.0, ref[L].abs().max().item()), (L, d)
        sp = feats.split[L]
        back = ((sp.hi.float() + sp.lo.float()) / ops.ACT_SCALE).permute(0, 3, 1, 2).cpu()
        assert (back - feats[L].cpu()).abs().max().item() < 1e-6 * max(1.0, ref[L].abs().max().item())
    if g is not None:
        assert np.abs(feats["layer4"].cpu().numpy() - g["ft_layer4"]).max() < 2e-4 * max(1.0, np.abs(g["ft_layer4"]).max())
    seg = GI.strip_prefix(case["seg"])
    logits_all = []
    for i, (P, Fw) in enumerate(case["PF"]):
        s_ref = F.conv2d(F.conv2d(ref["layer4"], P), Fw, None, 1, 1)
        cft = ops.conv2d_tc(feats.split["layer4"], ops.pack_conv_tc(P, device=DEV), out_f32=False, nchw=True)["nchw"]
        s = ops.corr3x3(cft, Fw.to(DEV)).unsqueeze(1)
        assert (s.cpu() - s_ref).abs().max() < 1e-4
        lg = trk.refiner(s, feats, img.shape[-2:])
        lg_ref = R.seg_forward(seg, s_ref, ref, img.shape[-2:])
        err = (lg.cpu() - lg_ref).abs().max().item()
        assert err < 1e-3, ("logits", i, err)                                   # north-star tolerance
        if g is not None:
            assert np.abs(lg.cpu().numpy() - g["logits%d" % i]).max() < 1e-3    # vs the executed reference
        logits_all.append((lg, lg_ref))
    # merged labels identical
    lut = torch.arange(n_obj + 1, dtype=torch.uint8)
    src = torch.cat([l[0][0] for l in logits_all], 0)
    masks, labels, _ = ops.merge_masks(src.contiguous(), (1 << n_obj) - 1, None, lut.to(DEV), False)
    cm = torch.zeros(n_obj + 1, *img.shape[-2:])
    for i, l in enumerate(logits_all):
        cm[i + 1] = torch.sigmoid(l[1][0, 0])
    merged_ref = R.merge_masks(cm)
    labels_ref = R.labels_from_masks(merged_ref.clone(), lut, False)
    # argmax label map: identical, except where the oracle's own decision flips under the +-1e-3 logit tolerance (with
    # random, unfitted (P, F) the objects overlap everywhere, so near-ties between saturated objects exist); every
    # mismatch must be such a tie and they must be rare.  The small fixture cases are exactly identical.
    import replay
    bad, tie = replay.tie_pixels(torch.stack([l[1][0, 0] for l in logits_all]), labels.cpu(), labels_ref.reshape(labels.shape), lut)
    print("P1 %s %s: %d mismatching label pixels (%d ties at +-1e-3) of %d" % (arch, tuple(size), bad, tie, labels.numel()))
    assert bad == tie and bad <= 1e-5 * labels.numel(), (bad, tie)
    if size == GI.MID:
        assert bad == 0
    # merged scores are softmax(p/(1-p)): where two objects saturate (p -> 1) the value is ill-conditioned by
    # construction (z = p/(1-p) ~ 1e3 amplifies a 1e-4 logit difference to O(0.1)), so compare robustly
    dm = (masks.cpu() - merged_ref).abs()
    assert (dm > 1e-3).float().mean().item() < 1e-3 and dm.mean().item() < 1e-4


def test_P1_batched_objects_and_frames_match_single():
    """The batched path (objects x frames in one pass) must give bit-identical logits to batch 1."""
    from frtm_vos_b200 import ops
    case = GI.feedforward_case("resnet18")
    trk, fe = _build("resnet18", case["bb"], case["seg"], GI.disc_params(256))
    from frtm_vos_b200 import synth
    seq = synth.SyntheticSequence(num_objects=2, num_frames=3, size=GI.MID, seq_id=9)
    imgs = torch.stack([seq[t][0] for t in range(3)]).to(DEV)
    nhwc, _, _ = fe.forward_split(imgs)
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(6, *nhwc["layer4"].hi.shape[1:3], generator=g).to(DEV)       # 3 frames x 2 objects
    lg = trk.refiner.forward_nhwc(scores, nhwc, GI.MID)
    for f in range(3):
        single, _, _ = fe.forward_split(imgs[f:f + 1])
        for n in range(2):
            one = trk.refiner.forward_nhwc(scores[f * 2 + n:f * 2 + n + 1], single, GI.MID)
            assert torch.equal(one[0], lg[f * 2 + n])


def _oracle_tracker(bb, seg, dp, hooks=None):
    from oracle import frtm_ref as R
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    return R.TrackerRef(bb, "resnet18", seg, GI.oracle_disc_params(dp), ImageAugmenter(GI.AUG_PARAMS).augment_first_frame,
                        "cpu", hooks=hooks)


def _e2e_setup(n_frames=18, size=GI.MID, n_obj=2):
    from frtm_vos_b200 import synth
    bb = synth.backbone_state_dict("resnet18", size=size)
    seg = synth.segnet_state_dict("resnet18")
    dp = GI.disc_params(256)
    seq = synth.SyntheticSequence(num_objects=n_obj, num_frames=n_frames, size=size, seq_id=3)
    return bb, seg, dp, seq, size


@pytest.mark.parametrize("size,n_obj", [(GI.MID, 2), (FULL, 3)])
def test_P4_end_to_end_oracle_replay(size, n_obj):
    """(128x224, 2 objects) and (480x854, 3 objects = BASELINE config 2's frame shape and object count), 18 frames.
    Inject each object's post-init (P, F, memory, CG state) from the oracle, then run the sequence: logits within
    1e-3 and identical label maps on every frame, through two GN updates per object.

    The 5-iteration CG update amplifies fp32 rounding differences of its inputs: on this small problem the reference
    itself moves its filter by 3e-5 .. 4e-4 (relative) when the memory samples are perturbed by 1e-7 .. 1e-6
    (measured with the oracle).  So each update is checked functionally against the oracle (relative filter error
    < 5e-3, the CUDA path lands at ~1e-4) and the oracle's filter is then re-injected, as SURVEY.md §4/P4 prescribes;
    the frame memory and CG state keep running free on the device."""
    from frtm_vos_b200 import ops
    bb, seg, dp, seq, size = _e2e_setup(size=size, n_obj=n_obj)
    dump = {"state": {}, "logits": {}}

    def after_init(oid, tm):
        dump["state"][oid] = dict(P=tm.P.clone(), F=tm.F.clone(), samples=tm.memory.samples.clone(),
                                  labels=tm.memory.labels.clone(), pw=tm.memory.pixel_weights.clone(),
                                  weights=tm.memory.weights.clone(), p=tm.optimizer.p[0].clone(),
                                  r_prev=tm.optimizer.r_prev[0].clone(), rho=tm.optimizer.rho.clone())

    def on_logits(frame, oid, s, lg):
        dump["logits"][(frame, oid)] = lg.clone()

    dump["F_upd"] = {}

    def after_update(frame, oid, tm):
        dump["F_upd"][(frame, oid)] = tm.F.clone()

    orc = _oracle_tracker(bb, seg, dp, hooks=dict(after_init=after_init, logits=on_logits, after_update=after_update))
    torch.manual_seed(11)
    out_ref, _ = orc.run_sequence(seq)

    trk, fe = _build("resnet18", bb, seg, dp)
    trk.graph_blocks = False                               # the spies below need the per-kernel path
    got_logits = {}
    orig_fwd = trk.refiner.forward_nhwc

    def spy(scores, feats, im_size):
        lg = orig_fwd(scores, feats, im_size)
        nF = feats["layer4"].hi.shape[0]                   # a block of frames goes through in one pass
        per_frame_lg = lg.view(nF, -1, *lg.shape[-2:])
        for f in range(nF):
            got_logits[trk.current_frame + f] = per_frame_lg[f].clone()
        return lg

    trk.refiner.forward_nhwc = spy
    orig_init = trk.initialize

    def init_and_inject(image, labels, new_objects):
        r = orig_init(image, labels, new_objects)
        for oid in new_objects:
            st, d = dump["state"][oid], trk.targets[oid].discriminator
            d.project.weight.data.copy_(st["P"]); d.filter.weight.data.copy_(st["F"])
            m = d.memory
            m.samples.copy_(st["samples"]); m.labels.copy_(st["labels"]); m.pixel_weights.copy_(st["pw"])
            m.weights.copy_(st["weights"])
            sten, uty = ops.build_stencil(m.pixel_weights[:5], m.labels[:5], m.samples.shape[-2:])
            m.stencil[:5] = sten; m.uty[:5] = uty
            m.refresh_split()
            d.update_optimizer.set_state(st["p"].to(DEV), st["r_prev"].to(DEV), float(st["rho"]))
        trk._stack = None
        return r

    trk.initialize = init_and_inject
    got_F = {}
    orig_block = trk._track_block

    def block_and_record(images):
        r = orig_block(images)
        last = trk.current_frame + len(images) - 1          # filter updates happen on the last frame of a block
        for oid, t in trk.targets.items():
            key = (last, oid)
            if key in dump["F_upd"]:
                got_F[key] = t.discriminator.filter.weight.detach().cpu().clone()
                t.discriminator.filter.weight.data.copy_(dump["F_upd"][key])        # teacher-force the next frames
        return r

    trk._track_block = block_and_record
    torch.manual_seed(11)
    out, fps = trk.run_sequence(seq)
    f_err = {k: (got_F[k] - v).abs().max().item() / v.abs().max().item() for k, v in dump["F_upd"].items()}
    worst, per_frame = 0.0, {}
    for (frame, oid), lg_ref in dump["logits"].items():
        lg = got_logits[frame][seq.obj_ids.index(oid)].cpu()
        e = (lg - lg_ref[0, 0]).abs().max().item()
        per_frame[frame] = max(per_frame.get(frame, 0.0), e)
        worst = max(worst, e)
    assert len(f_err) == 2 * n_obj and max(f_err.values()) < 5e-3, f_err
    assert worst < 1e-3, "logit err per frame: %s | filter rel err after updates: %s" % (
        " ".join("%d:%.1e" % kv for kv in sorted(per_frame.items())), f_err)
    # Label maps: identical, except that a pixel whose oracle decision flips under a +-1e-3 perturbation of the logits
    # (the stated logit tolerance) is a tie, not an error.  Such ties are counted and must stay rare.
    from oracle import frtm_ref as R
    lut = torch.tensor([0] + list(seq.obj_ids), dtype=torch.uint8)
    ties = 0
    for i, (a, b) in enumerate(zip(out, out_ref)):
        a, b = a.reshape(size).cpu(), b.reshape(size)
        if torch.equal(a, b):
            continue
        bad = a != b
        lg = torch.stack([dump["logits"][(i, oid)][0, 0] for oid in seq.obj_ids])
        explained = torch.zeros_like(bad)
        import itertools
        for signs in itertools.product((-1e-3, 1e-3), repeat=n_obj):
            cm = torch.zeros(n_obj + 1, *size)
            cm[1:] = torch.sigmoid(lg + torch.tensor(signs).view(n_obj, 1, 1))
            alt = R.labels_from_masks(R.merge_masks(cm), lut, False)
            explained |= (alt == a)
        assert bool((explained | ~bad).all()), "labels differ on frame %d beyond logit-tolerance ties" % i
        ties += int(bad.sum())
    # reported, not hidden: pixels whose label depends on the 1e-3 logit tolerance (of %d x %d x %d decided pixels)
    print("P4 %s: worst logit err %.2e, tie pixels %d of %d" % (size, worst, ties, len(out) * size[0] * size[1]))
    assert ties <= 8 * (size[0] * size[1]) // (GI.MID[0] * GI.MID[1]), ties
    for oid in seq.obj_ids:
        d, m = trk.targets[oid].discriminator, orc.targets[oid]["model"]
        assert torch.allclose(d.memory.weights.cpu(), m.memory.weights, atol=1e-6)
        assert d.memory.previous_replace_ind == m.memory.prev_ind and d.memory.current_size == m.memory.size


def test_P5_end_to_end_free_running():
    bb, seg, dp, seq, size = _e2e_setup()
    orc = _oracle_tracker(bb, seg, dp)
    torch.manual_seed(11)
    out_ref, _ = orc.run_sequence(seq)
    trk, fe = _build("resnet18", bb, seg, dp)
    torch.manual_seed(11)
    out, fps = trk.run_sequence(seq)
    a = torch.stack([o.reshape(size).cpu() for o in out])
    b = torch.stack([o.reshape(size) for o in out_ref])
    assert torch.equal(a[0], b[0])
    agree = (a == b).float().mean().item()
    # free-running: init is chaotic at ulp level, the reference agrees with itself (1 / 4 / 8 threads) on 99.992-99.998 % of
    # the pixels of this sequence (profiles/r02_oracle_spread.md); SURVEY.md §4 asks >= 99.9 % of the CUDA path
    print("P5 free-running label agreement with the oracle: %.6f (%d px differ)" % (agree, int((a != b).sum())))
    assert agree > 0.999, agree
    assert trk.targets[1].discriminator.memory.current_size == orc.targets[1]["model"].memory.size


def test_block_batching_is_exact():
    """Eight frames per pass must give exactly the frame-by-frame results (labels, filters, memory)."""
    bb, seg, dp, seq, size = _e2e_setup()
    res = []
    for batching in (False, True):
        trk, fe = _build("resnet18", bb, seg, dp)
        trk.block_batching = batching
        torch.manual_seed(11)
        out, _ = trk.run_sequence(seq)
        res.append((torch.stack([o.reshape(size).cpu() for o in out]),
                    [trk.targets[o].discriminator.filter.weight.detach().cpu().clone() for o in seq.obj_ids],
                    [trk.targets[o].discriminator.memory.weights.cpu().clone() for o in seq.obj_ids],
                    [trk.targets[o].discriminator.memory.samples.cpu().clone() for o in seq.obj_ids]))
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1] + res[0][2] + res[0][3], res[1][1] + res[1][2] + res[1][3]):
        assert torch.equal(a, b)


def test_no_cpu_fallback():
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200 import synth
    fe = ResnetFeatureExtractor("resnet18", state_dict=synth.backbone_state_dict("resnet18", size=GI.SMALL))
    with pytest.raises(RuntimeError):
        fe.to("cpu")
