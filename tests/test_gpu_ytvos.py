"""-m gpu: SURVEY.md §8 row f4 — the all-frames YouTubeVOS variant (``ytvos_validation/``): true-bicubic ``Upsampler``,
single-stage label rule over re-inserted ground truth, 'thresh' update labels, and the driver end to end against the oracle
restatement (``oracle.frtm_ref.YtvosTrackerRef``), whose own pieces are pinned to the executed reference by
``tests/golden/ytvos.npz``."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,size", [((2, 9, 13, 64), (18, 26)), ((1, 30, 54, 12), (70, 101)), ((3, 7, 5, 4), (7, 5)),
                                        ((1, 12, 17, 8), (5, 40))])
def test_resize_bicubic_matches_aten(shape, size):
    from frtm_vos_b200 import ops
    g = torch.Generator().manual_seed(shape[1] + size[0])
    x = torch.randn(*shape, generator=g)
    ref = F.interpolate(x.permute(0, 3, 1, 2), size, mode="bicubic", align_corners=False).permute(0, 2, 3, 1)
    y = ops.resize_bicubic(x.to(DEV), size)
    assert (y.cpu() - ref).abs().max().item() < 2e-6 * max(1.0, ref.abs().max().item())


def test_bicubic_upsampler_tail_matches_reference_fixture(golden):
    """``Upsampler`` = bicubic x2 -> conv1 + ReLU -> bicubic -> conv2, evaluated as in ``SegNetwork`` (conv2 contracted to 9
    tap maps in conv1's epilogue, resized, tap-shifted sum) against the executed ``ytvos_validation.seg_network.Upsampler``."""
    from frtm_vos_b200 import ops
    from frtm_vos_b200.model.seg_network import SegNetwork
    g = golden("ytvos")
    case = GI.ytvos_case()
    net = SegNetwork(1, 64, dict(layer5=512, layer4=256, layer3=128, layer2=64), True, upsampler="bicubic")
    sd = net.state_dict()
    for k, v in case["up"].items():
        sd[k] = v
    net.load_state_dict(sd)
    net.to(DEV)
    net._pack()
    P = net._packed
    x = case["x"].permute(0, 2, 3, 1).contiguous().to(DEV)
    h, w = x.shape[1:3]
    u = ops.split_f16(ops.resize_bicubic(x, (2 * h, 2 * w)))
    t12 = ops.conv2d_tc(u, P["up1"], relu=True, out_f32=False, tapw=P["up2_w"])["tap"]
    logits = ops.shift_sum9(ops.resize_bicubic(t12, case["image_size"]), P["up2_b"])
    assert np.abs(logits.cpu().numpy() - g["logits"][:, 0]).max() < 1e-4


def test_labels_from_probs_and_thresh_labels_match_fixture(golden):
    from frtm_vos_b200 import ops
    g = golden("ytvos")
    case = GI.ytvos_case()
    lut = torch.tensor([0, 3, 5, 9], dtype=torch.uint8, device=DEV)
    labels = ops.labels_from_probs(case["probs"].permute(1, 0, 2, 3).contiguous().to(DEV), lut)      # (frames, N, H, W)
    assert np.array_equal(labels.cpu().numpy(), g["labels"])                                           # bit-exact
    yb = ops.threshold(case["soft"].to(DEV), 0.5)
    assert np.array_equal(yb.cpu().numpy(), g["yb"])
    pw = ops.pixel_weights(yb, 0.1, True)
    assert np.abs(pw.cpu().numpy() - g["pw"]).max() < 1e-6
    s = ops.sigmoid_suppress(torch.tensor([[[0.0, 2.0], [-1.0, 30.0]]], device=DEV), torch.tensor([[0, 1], [0, 0]], dtype=torch.uint8, device=DEV))
    assert torch.allclose(s.cpu(), torch.tensor([[[0.5, 0.0], [0.26894143, 1.0]]]), atol=1e-6)


def test_forget_factor_zero_resets_cg_state():
    """direction_forget_factor = 0 (CG_forgetting_rate None in the variant, ``discriminator.py:254-257``): every run starts
    from a fresh CG state — two runs equal two runs of the oracle with forget 0."""
    from oracle import frtm_ref as R
    from frtm_vos_b200.model.discriminator import DiscriminatorLoss
    from frtm_vos_b200.model.optimizer import GaussNewtonCG
    from frtm_vos_b200.lib.tensorlist import TensorList
    from test_gpu_target_model import _fill_memory
    prob = GI.update_problem()
    mem = _fill_memory(prob)
    filt = prob["F0"].clone().to(DEV)
    problem = DiscriminatorLoss(x=mem.samples, y=mem.labels, filter_regs=(1e-2,), precond=(1e-2,), sample_weights=mem.weights,
                                net=None, pixel_weighting=mem.pixel_weights, memory=mem)
    opt = GaussNewtonCG(problem, TensorList([filt]), fletcher_reeves=False, standard_alpha=True, direction_forget_factor=0)
    om = R.FrameMemory(prob["samples"].shape[0], prob["samples"].shape[1:], prob["labels"].shape[1:], "cpu", 0.1)
    om.samples.copy_(prob["samples"]); om.labels.copy_(prob["labels"]); om.pixel_weights.copy_(prob["pixel_weights"])
    om.weights.copy_(prob["weights"])
    Fo = prob["F0"].clone()
    oopt = R.GaussNewtonCGRef(R.GNProblem(om, (1e-2,), (1e-2,), False), [Fo], 0.0)
    for _ in range(2):
        opt.run((5,))
        oopt.run((5,))
        assert (filt.cpu() - Fo).abs().max().item() < 1e-5 * max(1.0, Fo.abs().max().item())


def test_ytvos_tracker_against_oracle():
    """The all-frames driver end to end (objects starting on frames 0 and 3) against the oracle restatement: identical labels
    on the frames fixed by construction (ground truth re-inserted), >= 99.5 % agreement free-running (init is chaotic at ulp
    level, profiles/r02_oracle_spread.md), every object present in the output."""
    from oracle import frtm_ref as R
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200.model.seg_network import SegNetwork
    from frtm_vos_b200.model.ytvos import YtvosTracker
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    size = GI.MID
    bb = synth.backbone_state_dict("resnet18", size=size)
    seg = synth.segnet_state_dict("resnet18")
    dp = GI.disc_params(256)
    dp["pixel_weighting"] = dict(method="hinge", tf=0.1, update_method="thresh")
    seq = synth.SyntheticSequence(num_objects=2, num_frames=13, size=size, seq_id=5, start_frames=[0, 3])
    fe = ResnetFeatureExtractor("resnet18", state_dict=bb).to(DEV)
    refiner = SegNetwork(1, 64, {L: c for L, c in fe.get_out_channels().items() if L != "layer1"}, True, upsampler="bicubic")
    trk = YtvosTracker(ImageAugmenter(GI.AUG_PARAMS), fe, dict(dp, device=DEV), refiner, DEV)
    trk.load_state_dict(seg)
    trk.to(DEV)
    torch.manual_seed(11)
    out, fps = trk.run_sequence(seq)
    odp = GI.oracle_disc_params(dp)
    orc = R.YtvosTrackerRef(bb, "resnet18", seg, odp, ImageAugmenter(GI.AUG_PARAMS).augment_first_frame, "cpu")
    torch.manual_seed(11)
    out_ref, _ = orc.run_sequence(seq)
    a = torch.stack([o.reshape(size).cpu() for o in out])
    b = torch.stack([o.reshape(size) for o in out_ref])
    assert len(out) == 13 and fps > 0
    assert torch.equal(a[0], b[0])                                             # frame 0 = ground truth of object 1
    gt3 = seq[3][1][0]
    assert bool(((a[3] == 2) == (gt3 == 2)).all())                            # object 2's first frame carries its ground truth
    agree = (a == b).float().mean().item()
    print("ytvos free-running label agreement with the oracle: %.6f" % agree)
    assert agree > 0.995, agree
    assert set(a.unique().tolist()) == {0, 1, 2}
