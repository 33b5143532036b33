"""ORACLE tooling — generate ``tests/golden/*.npz`` by executing the REAL reference (build container only).

    python -m oracle.make_golden            # writes tests/golden/, asserts the restatement agrees

The reference has no tests or golden vectors of its own, so these fixtures — outputs of the
reference's own classes on the seeded inputs of ``tests/golden_inputs.py`` — are the pin for
``oracle/frtm_ref.py`` and, through it, for the CUDA path.  Fixtures: ``update`` (GaussNewtonCG on the filter-only
problem, two consecutive runs so the persistent CG state is exercised), ``init_step`` (RHS and one J^T J product
of the joint project/filter problem at a fixed point), ``memory`` (sample-weight / replace-index trace),
``pixel_weights``, ``merge``, ``feedforward`` (backbone -> apply -> refinement logits at fixed (P, F)),
``e2e`` (free-running label maps + target-model state of a short two-object sequence), ``eval`` (J / F measures,
sequence statistics and report bar graphs of lib/davis.py / lib/utils.py on seeded label maps).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import shims, frtm_ref as R  # noqa: E402
import golden_inputs as GI  # noqa: E402
from frtm_vos_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _ref_memory(ref, prob):
    cap = prob["samples"].shape[0]
    mem = ref.memory.Memory(cap, prob["samples"].shape[1:], prob["labels"].shape[1:], "cpu", 0.1)
    mem.samples.copy_(prob["samples"]); mem.labels.copy_(prob["labels"])
    mem.pixel_weights.copy_(prob["pixel_weights"]); mem.weights.copy_(prob["weights"])
    mem.current_size = int((prob["weights"] > 0).sum())
    return mem


def gen_update(ref):
    prob = GI.update_problem()
    mem = _ref_memory(ref, prob)
    filt = ref.discriminator.conv(96, 1, 3, bias=False)
    with torch.no_grad():
        filt.weight.copy_(prob["F0"])
    TL = ref.tensorlist.TensorList
    problem = ref.discriminator.DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-2,), precond=(1e-2,),
                                                  sample_weights=mem.weights, net=filt, pixel_weighting=mem.pixel_weights)
    opt = ref.optimizer.GaussNewtonCG(problem, TL([filt.weight]), fletcher_reeves=False, standard_alpha=True,
                                      direction_forget_factor=(1 - 0.1) ** 750)
    opt.run((10,))
    F1 = filt.weight.detach().clone()
    p1, rho1, rprev1 = opt.p[0].clone(), opt.rho[0].clone(), opt.r_prev[0].clone()
    b1 = opt.b[0].clone()
    # perturb the memory like an insert would and run again (persistent p / rho / r_prev in play)
    g = torch.Generator().manual_seed(77)
    mem.samples[3] = torch.randn(mem.samples[3].shape, generator=g) * 0.5
    mem.labels[3] = GI.blob_masks(1, mem.labels.shape[-2:], g, soft=True)[0]
    opt.run((10,))
    F2 = filt.weight.detach().clone()
    # restatement must agree
    om = R.FrameMemory(prob["samples"].shape[0], prob["samples"].shape[1:], prob["labels"].shape[1:], "cpu", 0.1)
    om.samples.copy_(prob["samples"]); om.labels.copy_(prob["labels"]); om.pixel_weights.copy_(prob["pixel_weights"])
    om.weights.copy_(prob["weights"])
    Fo = prob["F0"].clone()
    oo = R.GaussNewtonCGRef(R.GNProblem(om, (1e-2,), (1e-2,), False), [Fo], (1 - 0.1) ** 750)
    oo.run((10,))
    assert torch.equal(Fo, F1), (Fo - F1).abs().max()
    om.samples[3] = mem.samples[3]; om.labels[3] = mem.labels[3]
    oo.run((10,))
    assert torch.equal(Fo, F2), (Fo - F2).abs().max()
    np.savez_compressed(os.path.join(OUT, "update.npz"), F1=_np(F1), F2=_np(F2), b1=_np(b1), p1=_np(p1), rho1=_np(rho1),
                        rprev1=_np(rprev1), s3=_np(mem.samples[3]), l3=_np(mem.labels[3]))
    print("update: |F1-F0| %.3e |F2-F1| %.3e" % ((F1 - prob["F0"]).abs().max(), (F2 - F1).abs().max()))


def gen_init_step(ref):
    ip = GI.init_problem()
    K = ip["x"].shape[0]
    dsc = ref.discriminator.Discriminator(in_channels=ip["x"].shape[1], pixel_weighting=dict(method="hinge", tf=0.1),
                                          device="cpu")
    pw = dsc.compute_pixel_weights(ip["y"])
    mem = ref.memory.Memory(K, ip["x"].shape[-3:], ip["y"].shape[-3:], "cpu", 0.1)
    mem.initialize(ip["x"], ip["y"], pw)
    with torch.no_grad():
        dsc.project.weight.copy_(ip["P0"]); dsc.filter.weight.copy_(ip["F0"])
    TL = ref.tensorlist.TensorList
    theta = TL([dsc.project.weight, dsc.filter.weight])
    problem = ref.discriminator.DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-4, 1e-2),
                                                  precond=(1e-4, 1e-2), sample_weights=mem.weights,
                                                  net=torch.nn.Sequential(dsc.project, dsc.filter),
                                                  pixel_weighting=mem.pixel_weights)
    opt = ref.optimizer.GaussNewtonCG(problem, theta, fletcher_reeves=False, standard_alpha=True,
                                      direction_forget_factor=(1 - 0.1) ** 750)
    # one GN linearisation by hand: RHS and a J^T J product (optimizer.py:77-85,155-157)
    problem.initialize()
    theta.requires_grad_(True)
    opt.f0 = problem(theta)
    opt.g = opt.f0.detach(); opt.g.requires_grad_(True)
    opt.dfdxt_g = TL(torch.autograd.grad(opt.f0, theta, opt.g, create_graph=True))
    b = -opt.dfdxt_g.detach()
    Ap = opt.A(TL([ip["dP"], ip["dF"]]))
    theta.detach_()
    # now the free-running init, short schedule
    dsc2 = ref.discriminator.Discriminator(in_channels=ip["x"].shape[1], init_iters=(5, 10), update_iters=(5,),
                                           memory_size=8, CG_forgetting_rate=750,
                                           pixel_weighting=dict(method="hinge", tf=0.1), device="cpu")
    with torch.no_grad():
        dsc2.project.weight.copy_(ip["P0"]); dsc2.filter.weight.copy_(ip["F0"])
    dsc2.init(ip["x"], ip["y"].byte())
    s_fin = dsc2(ip["x"]).detach()
    # restatement must agree bit for bit
    tm = R.TargetModelRef(ip["x"].shape[1], init_iters=(5, 10), update_iters=(5,), memory_size=8,
                          seed_weights=(ip["P0"].clone(), ip["F0"].clone()))
    tm.init(ip["x"], ip["y"].byte())
    assert torch.equal(tm.P, dsc2.project.weight) and torch.equal(tm.F, dsc2.filter.weight)
    np.savez_compressed(os.path.join(OUT, "init_step.npz"), pw=_np(pw), bP=_np(b[0]), bF=_np(b[1]), AP=_np(Ap[0]),
                        AF=_np(Ap[1]), P_fin=_np(dsc2.project.weight), F_fin=_np(dsc2.filter.weight), s_fin=_np(s_fin),
                        w_fin=_np(dsc2.memory.weights))
    print("init_step: |b_P| %.3e |A dP| %.3e score range %.3f..%.3f" % (b[0].abs().max(), Ap[0].abs().max(),
                                                                          s_fin.min(), s_fin.max()))


def gen_memory(ref):
    mem = ref.memory.Memory(10, (1, 1, 1), (1, 1, 1), "cpu", 0.1)
    z = torch.zeros(5, 1, 1, 1)
    mem.initialize(z, z, z)
    trace_w, trace_r = [_np(mem.weights).copy()], []
    om = R.FrameMemory(10, (1, 1, 1), (1, 1, 1), "cpu", 0.1)
    om.fill(z, z, z)
    for i in range(25):
        mem.update(z[0], z[0], z[0])
        om.insert(z[0], z[0], z[0])
        assert torch.equal(mem.weights, om.weights) and mem.previous_replace_ind == om.prev_ind
        trace_w.append(_np(mem.weights).copy()); trace_r.append(mem.previous_replace_ind)
    np.savez_compressed(os.path.join(OUT, "memory.npz"), weights=np.stack(trace_w), replace=np.array(trace_r))
    print("memory: replace trace", trace_r)


def gen_pixel_weights_merge(ref):
    g = torch.Generator().manual_seed(21)
    y = GI.blob_masks(6, GI.SMALL, g, soft=False)
    y[4] = 0
    y[4, 0, 0, :5] = 1          # < 10 px  -> "too small" branch
    y[5] = (GI.blob_masks(1, GI.SMALL, g)[0] * 0 + 1)   # everything foreground
    y[5, 0, :8] = 0
    dsc = ref.discriminator.Discriminator(in_channels=8, pixel_weighting=dict(method="hinge", tf=0.1), device="cpu")
    pw = dsc.compute_pixel_weights(y)
    assert torch.equal(pw, R.pixel_weights(y, 0.1))
    # merge: run the reference's Tracker.track tail on a synthetic current_masks via the restated function,
    # and the real code path via a minimal stand-in (tracker.py:214-221 is inline code, so execute it verbatim-in-effect)
    probs = torch.rand(4, *GI.SMALL, generator=g)
    probs[0] = 0
    probs[2, :10] = 1.0   # saturated -> clamp
    probs[3, -10:] = 0.0
    trk = ref.tracker.Tracker.__new__(ref.tracker.Tracker)
    torch.nn.Module.__init__(trk)
    trk.targets = {}
    trk.current_masks = probs.clone()
    trk.current_frame = 1
    trk.disc_params = shims._AttrDict(update_filters=False)
    trk.feature_extractor = lambda image: {}
    merged = trk.track(torch.zeros(3, *GI.SMALL))
    assert torch.equal(merged, R.merge_masks(probs.clone()))
    lut = torch.tensor([0, 3, 5, 9], dtype=torch.uint8)
    labels = R.labels_from_masks(merged.clone(), lut, False)
    np.savez_compressed(os.path.join(OUT, "pw_merge.npz"), pw=_np(pw), merged=_np(merged), labels=_np(labels))
    print("pixel_weights/merge ok; label histogram", torch.bincount(labels.flatten().long()).tolist())


def _ref_tracker(ref, arch, bb, seg, dp):
    ED = ref.EasyDict
    fe = ref.feature_extractor.ResnetFeatureExtractor(arch).to("cpu")
    fe.resnet.load_state_dict(bb, strict=False)
    chans = fe.get_out_channels()
    refiner = ref.seg_network.SegNetwork(1, 64, {L: c for L, c in chans.items()
                                                 if L in ("layer5", "layer4", "layer3", "layer2")}, True)
    trk = ref.tracker.Tracker(ref.augmenter.ImageAugmenter(ED(GI.AUG_PARAMS)), fe, ED(dp), refiner, "cpu")
    trk.load_state_dict(seg)
    trk.eval()
    return trk, fe


def gen_feedforward(ref, arch="resnet18"):
    case = GI.feedforward_case(arch)
    C = synth.backbone_out_channels(arch)["layer4"]
    trk, fe = _ref_tracker(ref, arch, case["bb"], case["seg"], GI.disc_params(C))
    img = case["image"]
    with torch.no_grad():
        feats = fe(img)
        out = dict()
        for L in ("layer1", "layer2", "layer3", "layer4", "layer5"):
            out["ft_" + L] = _np(feats[L]) if L in ("layer4", "layer5") else _np(feats[L][:, :8])
        logits = []
        for i, (P, Fw) in enumerate(case["PF"]):
            s = torch.nn.functional.conv2d(torch.nn.functional.conv2d(feats["layer4"], P), Fw, None, 1, 1)
            lg = trk.refiner(s, feats, img.shape[-2:])
            out["scores%d" % i] = _np(s)
            logits.append(lg)
            out["logits%d" % i] = _np(lg)
        # restatement
        of = R.backbone_features(case["bb"], arch, img)
        for L in of:
            assert torch.equal(of[L], feats[L]), L
        for i, (P, Fw) in enumerate(case["PF"]):
            s = torch.nn.functional.conv2d(torch.nn.functional.conv2d(of["layer4"], P), Fw, None, 1, 1)
            lg = R.seg_forward(GI.strip_prefix(case["seg"]), s, of, img.shape[-2:])
            assert torch.equal(lg, logits[i])
    np.savez_compressed(os.path.join(OUT, "feedforward_%s.npz" % arch), **out)
    print("feedforward %s: logits range %.2f..%.2f" % (arch, min(l.min() for l in logits), max(l.max() for l in logits)))


def gen_e2e(ref):
    arch, size = "resnet18", GI.MID
    bb = synth.backbone_state_dict(arch, size=size)
    seg = synth.segnet_state_dict(arch)
    dp = GI.disc_params(256)
    seq = synth.SyntheticSequence(num_objects=2, num_frames=18, size=size, seq_id=3)
    trk, _ = _ref_tracker(ref, arch, bb, seg, dp)
    torch.manual_seed(11)
    out_ref, _ = trk.run_sequence(seq)
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    orc = R.TrackerRef(bb, arch, seg, GI.oracle_disc_params(dp), ImageAugmenter(GI.AUG_PARAMS).augment_first_frame)
    torch.manual_seed(11)
    out_or, _ = orc.run_sequence(seq)
    dump = dict(labels=np.stack([_np(o.reshape(size)) for o in out_ref]))
    for oid in seq.obj_ids:
        d = trk.targets[oid].discriminator
        m = orc.targets[oid]["model"]
        assert torch.equal(d.project.weight, m.P) and torch.equal(d.filter.weight, m.F)
        assert torch.equal(d.memory.weights, m.memory.weights)
        dump["P%d" % oid], dump["F%d" % oid], dump["w%d" % oid] = _np(d.project.weight), _np(d.filter.weight), _np(d.memory.weights)
    for a, b in zip(out_ref, out_or):
        assert torch.equal(a.reshape(size), b.reshape(size))
    np.savez_compressed(os.path.join(OUT, "e2e_rn18.npz"), **dump)
    print("e2e: %d frames, restatement bit-identical to the executed reference" % len(out_ref))


def gen_eval(ref):
    """J / F measures, sequence statistics and the bar graph of the executed reference (lib/davis.py, lib/utils.py:9-22) on the
    seeded maps of ``golden_inputs.eval_mask_pairs`` — the pin for ``frtm_vos_b200/lib/davis.py`` where the reference is absent."""
    pairs = GI.eval_mask_pairs()
    with ref.numpy1_aliases():
        J = np.array([float(ref.davis.davis_jaccard_measure(a.copy(), b.copy())) for a, b in pairs])
        F = np.array([float(ref.davis.davis_f_measure(a.copy(), b.copy())) for a, b in pairs])
        vecs = GI.eval_score_vectors()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            stats = np.array([[float(ref.davis.mean(v)), float(ref.davis.recall(v)), float(ref.davis.decay(v)), float(ref.davis.std(v))]
                              for v in vecs])
        bars = np.array([ref.utils.text_bargraph(np.concatenate((v, [-0.1, 1.2]))) for v in vecs])
    np.savez_compressed(os.path.join(OUT, "eval.npz"), J=J, F=F, stats=stats, bars=bars)
    print("eval: %d mask pairs (J mean %.3f, F mean %.3f), %d score vectors" % (len(pairs), J.mean(), F.mean(), len(vecs)))


def gen_ytvos(ref):
    """Pieces of the all-frames YouTubeVOS variant (ytvos_validation/): the bicubic ``Upsampler``, ``merge_segmentations``
    and the hinge weights / binary labels of the 'thresh' update method, executed from the reference's own modules."""
    import importlib
    sn = importlib.import_module("ytvos_validation.seg_network")
    tr = importlib.import_module("ytvos_validation.tracker")
    dc = importlib.import_module("ytvos_validation.discriminator")
    case = GI.ytvos_case()
    up = sn.Upsampler(64)
    up.load_state_dict({k[len("project."):]: v for k, v in case["up"].items()})
    with torch.no_grad():
        logits = up(case["x"], case["image_size"])
    mine = R.upsampler_bicubic(case["up"], case["x"], case["image_size"])
    assert torch.equal(logits, mine), "Upsampler restatement differs from the reference"
    segs, ids = tr.Tracker.merge_segmentations(case["probs"], [3, 5, 9])
    mine = R.merge_segmentations(case["probs"])
    assert torch.equal(segs, mine) and ids.tolist() == [0, 3, 5, 9]
    labels = ids[segs.argmax(dim=0)]
    # 'thresh' update: binary labels + per-frame hinge weights (discriminator.py:159-215,364-367)
    E = ref.EasyDict
    params = E(layer="layer4", cdims=96, kernel_size=[3], filter_reg=[1e-4, 1e-2], precon=[1e-4, 1e-2], n_channels=[1],
               with_bias=False, init_iters=[5], update_iters=[5],
               pixel_weighting=dict(method="hinge", tf=0.1, distractor_mult=1.0, per_frame=True, update_method="thresh",
                                    max_fg_weight=100))
    d = dc.Discriminator(params)
    pw, yb = d.get_online_weights(case["soft"])
    mine_y = (case["soft"] > 0.5).float()
    assert torch.equal(yb, mine_y)
    mine_pw = R.pixel_weights(mine_y, 0.1)
    assert torch.equal(pw, mine_pw), "thresh-mode hinge weights differ from the restatement"
    np.savez_compressed(os.path.join(OUT, "ytvos.npz"), logits=_np(logits), segs=_np(segs), labels=_np(labels).astype(np.uint8),
                        pw=_np(pw), yb=_np(yb))
    print("ytvos: Upsampler %s, merge_segmentations %s, thresh weights %s — restatement bit-identical"
          % (tuple(logits.shape), tuple(segs.shape), tuple(pw.shape)))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = shims.load_reference()
    gen_memory(ref)
    gen_pixel_weights_merge(ref)
    gen_update(ref)
    gen_init_step(ref)
    gen_feedforward(ref, "resnet18")
    gen_feedforward(ref, "resnet101")
    gen_e2e(ref)
    gen_eval(ref)
    gen_ytvos(ref)


if __name__ == "__main__":
    if len(sys.argv) > 1:            # python -m oracle.make_golden eval  -> only the named fixtures
        os.makedirs(OUT, exist_ok=True)
        _ref = shims.load_reference()
        for _name in sys.argv[1:]:
            globals()["gen_" + _name](_ref)
    else:
        main()
