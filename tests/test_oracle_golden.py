"""The oracle restatement (oracle/frtm_ref.py) against fixtures produced by the EXECUTED reference
(oracle/make_golden.py).  In the build container the agreement is bit-exact (asserted at generation time);
here tolerances are a few ulp so the suite also passes on hosts whose CPU kernels round differently."""
import numpy as np
import torch

import golden_inputs as GI
from oracle import frtm_ref as R


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_memory_trace(golden):
    g = golden("memory")
    z = torch.zeros(5, 1, 1, 1)
    m = R.FrameMemory(10, (1, 1, 1), (1, 1, 1), "cpu", 0.1)
    m.fill(z, z, z)
    assert np.allclose(m.weights.numpy(), g["weights"][0], atol=1e-7)
    for i in range(25):
        m.insert(z[0], z[0], z[0])
        assert m.prev_ind == int(g["replace"][i])
        assert np.allclose(m.weights.numpy(), g["weights"][i + 1], atol=1e-7)


def test_pixel_weights_and_merge(golden):
    g = golden("pw_merge")
    gen = torch.Generator().manual_seed(21)
    y = GI.blob_masks(6, GI.SMALL, gen, soft=False)
    y[4] = 0
    y[4, 0, 0, :5] = 1
    y[5] = (GI.blob_masks(1, GI.SMALL, gen)[0] * 0 + 1)
    y[5, 0, :8] = 0
    assert np.allclose(R.pixel_weights(y, 0.1).numpy(), g["pw"], atol=1e-6)
    probs = torch.rand(4, *GI.SMALL, generator=gen)
    probs[0] = 0
    probs[2, :10] = 1.0
    probs[3, -10:] = 0.0
    merged = R.merge_masks(probs)
    assert np.allclose(merged.numpy(), g["merged"], atol=1e-6)
    lut = torch.tensor([0, 3, 5, 9], dtype=torch.uint8)
    assert np.array_equal(R.labels_from_masks(merged.clone(), lut, False).numpy(), g["labels"])


def test_update_phase_two_runs(golden):
    g = golden("update")
    prob = GI.update_problem()
    om = R.FrameMemory(prob["samples"].shape[0], prob["samples"].shape[1:], prob["labels"].shape[1:], "cpu", 0.1)
    om.samples.copy_(prob["samples"]); om.labels.copy_(prob["labels"])
    om.pixel_weights.copy_(prob["pixel_weights"]); om.weights.copy_(prob["weights"])
    Fo = prob["F0"].clone()
    opt = R.GaussNewtonCGRef(R.GNProblem(om, (1e-2,), (1e-2,), False), [Fo], (1 - 0.1) ** 750)
    opt.run((10,))
    assert np.allclose(Fo.numpy(), g["F1"], atol=2e-6)
    assert np.allclose(opt.p[0].numpy(), g["p1"], rtol=1e-3, atol=1e-5 * np.abs(g["p1"]).max())
    om.samples[3] = _t(g["s3"]); om.labels[3] = _t(g["l3"])
    opt.run((10,))
    assert np.allclose(Fo.numpy(), g["F2"], atol=2e-6)


def test_init_free_running(golden):
    g = golden("init_step")
    ip = GI.init_problem()
    assert np.allclose(R.pixel_weights(ip["y"], 0.1).numpy(), g["pw"], atol=1e-6)
    tm = R.TargetModelRef(ip["x"].shape[1], init_iters=(5, 10), update_iters=(5,), memory_size=8,
                          seed_weights=(ip["P0"].clone(), ip["F0"].clone()))
    tm.init(ip["x"], ip["y"].byte())
    s = torch.nn.functional.conv2d(torch.nn.functional.conv2d(ip["x"], tm.P), tm.F, None, 1, 1)
    # the init phase amplifies ulp-level differences (SURVEY.md finding 8) -> functional tolerance off-box
    assert np.abs(s.numpy() - g["s_fin"]).max() < 5e-2
    assert np.allclose(tm.memory.weights.numpy(), g["w_fin"], atol=1e-7)


import pytest


@pytest.mark.parametrize("arch", ["resnet18", "resnet101"])
def test_feedforward_fixed_state(golden, arch):
    g = golden("feedforward_" + arch)
    case = GI.feedforward_case(arch)
    feats = R.backbone_features(case["bb"], arch, case["image"])
    for L in ("layer4", "layer5"):
        assert np.allclose(feats[L].numpy(), g["ft_" + L], atol=1e-4), L
    for L in ("layer1", "layer2", "layer3"):
        assert np.allclose(feats[L][:, :8].numpy(), g["ft_" + L], atol=1e-4), L
    seg = GI.strip_prefix(case["seg"])
    for i, (P, Fw) in enumerate(case["PF"]):
        s = torch.nn.functional.conv2d(torch.nn.functional.conv2d(feats["layer4"], P), Fw, None, 1, 1)
        assert np.allclose(s.numpy(), g["scores%d" % i], atol=1e-4)
        lg = R.seg_forward(seg, s, feats, case["image"].shape[-2:])
        assert np.abs(lg.numpy() - g["logits%d" % i]).max() < 1e-3


def test_e2e_labels(golden):
    g = golden("e2e_rn18")
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    size = GI.MID
    bb = synth.backbone_state_dict("resnet18", size=size)
    seg = synth.segnet_state_dict("resnet18")
    seq = synth.SyntheticSequence(num_objects=2, num_frames=18, size=size, seq_id=3)
    orc = R.TrackerRef(bb, "resnet18", seg, GI.oracle_disc_params(GI.disc_params(256)),
                       ImageAugmenter(GI.AUG_PARAMS).augment_first_frame)
    torch.manual_seed(11)
    out, _ = orc.run_sequence(seq)
    lab = np.stack([o.reshape(size).numpy() for o in out])
    agree = (lab == g["labels"]).mean()
    assert np.array_equal(lab[0], g["labels"][0])
    assert agree > 0.98, agree    # bit-identical in the build container; chaotic init elsewhere (finding 8)


def test_ytvos_variant_pieces_match_reference_fixture(golden):
    """Row f4: the all-frames variant's own pieces (bicubic Upsampler, merge_segmentations + labels, 'thresh' update labels
    and hinge weights) of the restatement against the fixture produced by executing ytvos_validation/*.py."""
    from oracle import frtm_ref as R
    g = golden("ytvos")
    case = GI.ytvos_case()
    logits = R.upsampler_bicubic(case["up"], case["x"], case["image_size"])
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-5
    segs = R.merge_segmentations(case["probs"])
    assert np.abs(segs.numpy() - g["segs"]).max() < 1e-6
    lut = torch.tensor([0, 3, 5, 9], dtype=torch.uint8)
    assert np.array_equal(lut[segs.argmax(dim=0)].numpy(), g["labels"])
    yb = (case["soft"] > 0.5).float()
    assert np.array_equal(yb.numpy(), g["yb"])
    assert np.abs(R.pixel_weights(yb, 0.1).numpy() - g["pw"]).max() < 1e-6
