"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, host-side containers and
data generators behave, and (build container only) the host-side augmenter reproduces the reference's."""
import ctypes
import os

import numpy as np
import pytest
import torch

import golden_inputs as GI


def test_library_exports_every_declared_symbol():
    from frtm_vos_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 25
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), "libfrtm_b200.so does not export %s" % name
    L = _lib.lib()
    assert L.version() >= 100
    assert L.launch_count() == 0            # nothing may have launched on a CPU-only host


def test_header_cites_reference_and_is_plain_c():
    src = open(os.path.join(os.path.dirname(__file__), "..", "include", "frtm_b200.h")).read()
    assert 'extern "C"' in src and "torch" not in src.replace("pytorch", "").lower().replace("torchvision", "")
    for cite in ("model/discriminator.py", "model/optimizer.py", "model/memory.py", "model/tracker.py",
                 "model/seg_network.py", "model/feature_extractor.py", "nppig.cpp"):
        assert cite in src, cite


def test_product_fails_loudly_without_cuda():
    from frtm_vos_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU host")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.normalize_u8(torch.zeros(1, 3, 8, 8, dtype=torch.uint8))
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200 import synth
    fe = ResnetFeatureExtractor("resnet18", state_dict=synth.backbone_state_dict("resnet18", size=GI.SMALL))
    with pytest.raises(RuntimeError, match="CUDA"):
        fe.to("cpu")


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import frtm_vos_b200.model.tracker, frtm_vos_b200.model.seg_network, frtm_vos_b200.ops; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root)


def test_tensorlist_semantics():
    from frtm_vos_b200.lib.tensorlist import TensorList
    a = TensorList([torch.arange(3.0) + 1, torch.ones(2) * 2])
    b = TensorList([torch.ones(3) * 3, torch.ones(2) * 5])
    assert torch.equal((a * b)[0], torch.tensor([3.0, 6.0, 9.0]))
    assert torch.equal((2.0 - a)[1], torch.zeros(2))
    assert float(sum(a.view(-1) @ b.view(-1))) == 38.0
    assert isinstance(a[0:1], TensorList) and torch.is_tensor(a[0])
    c = a.clone()
    c += b
    c /= 2
    assert torch.equal(c[1], torch.ones(2) * 3.5) and torch.equal(a[1], torch.ones(2) * 2)
    assert not hasattr(a, "__torch_function__")
    with pytest.raises(AttributeError):
        a.not_a_tensor_method
    assert torch.equal((-a)[0], -(torch.arange(3.0) + 1))
    assert len(a.concat(b)) == 4 and len(TensorList([a, b]).unroll()) == 4


def test_synthetic_sequence_protocol():
    from frtm_vos_b200 import synth
    seq = synth.SyntheticSequence(num_objects=3, num_frames=5, size=GI.SMALL, seq_id=2)
    assert len(seq) == 5 and seq.obj_ids == [1, 2, 3] and len(seq.frame_names) == 5
    im, lb, new = seq[0]
    assert im.dtype == torch.uint8 and im.shape == (3, *GI.SMALL) and lb.shape == (1, *GI.SMALL) and new == [1, 2, 3]
    assert set(lb.unique().tolist()) == {0, 1, 2, 3}
    im1, lb1, new1 = seq[1]
    assert lb1 == [] and new1 == []
    again = synth.SyntheticSequence(num_objects=3, num_frames=5, size=GI.SMALL, seq_id=2)
    assert torch.equal(again[3][0], seq[3][0])
    late = synth.SyntheticSequence(num_objects=2, num_frames=5, size=GI.SMALL, seq_id=2, start_frames=[0, 2])
    assert late[0][2] == [1] and late[2][2] == [2] and set(late[2][1].unique().tolist()) <= {0, 2}
    sd = synth.segnet_state_dict("resnet18")
    assert len(sd) == 140 and sd["refiner.TSE.layer4.reduce.0.weight"].shape[1] == 256


def test_conv_weight_packing_layout():
    from frtm_vos_b200 import ops
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    pc = ops.pack_conv(w, torch.tensor([1.0, 2.0]), device="cpu")
    assert pc.w.shape == (3, 3 * 4, 4) and pc.cin_pad == 4
    assert pc.w[1, 2 * 4 + 1, 1] == w[1, 1, 1, 2] and pc.w[0, 3, 0] == 0
    bn = dict(weight=torch.tensor([2.0, 1.0]), bias=torch.tensor([0.5, 0.0]), running_mean=torch.tensor([1.0, 0.0]),
              running_var=torch.tensor([4.0, 1.0]))
    pc2 = ops.pack_conv(w, None, bn=bn, device="cpu", eps=0.0)
    assert torch.allclose(pc2.w[0, 0, 0], w[0, 0, 0, 0] * 1.0) and torch.allclose(pc2.bias, torch.tensor([-0.5, 0.0]))


@pytest.mark.needs_reference
def test_augmenter_matches_reference():
    from oracle import shims
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    ref = shims.load_reference()
    seq = synth.SyntheticSequence(num_objects=2, num_frames=1, size=GI.MID, seq_id=5)
    im, lb, _ = seq[0]
    mask = (lb == 2).byte()
    np.random.seed(0); torch.manual_seed(0)
    a_im, a_lb = ref.augmenter.ImageAugmenter(ref.EasyDict(GI.AUG_PARAMS)).augment_first_frame(im, mask)
    np.random.seed(0); torch.manual_seed(0)
    b_im, b_lb = ImageAugmenter(GI.AUG_PARAMS).augment_first_frame(im, mask)
    assert torch.equal(a_im, b_im) and torch.equal(a_lb, b_lb) and a_im.shape[0] == 5


def test_cropped_telea_inpaint_is_identical_to_full_frame():
    """The augmenter inpaints only the hole's bounding box (+ margin); the result must equal cv2.inpaint on the full frame."""
    import cv2
    import numpy as np
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import telea_inpaint_cropped
    seq = synth.SyntheticSequence(num_objects=3, num_frames=1, size=(240, 427), seq_id=3)
    im, lb, ids = seq[0]
    image = np.ascontiguousarray(im.numpy().transpose(1, 2, 0))
    for oid in ids:
        m = (lb[0].numpy() == oid).astype(np.uint8)[..., None]
        outer = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2)))
        for radius in (1, 3):
            full = cv2.inpaint(image, outer, inpaintRadius=radius, flags=cv2.INPAINT_TELEA)
            assert np.array_equal(telea_inpaint_cropped(image, outer, radius), full)
    # hole touching the frame border, and an empty hole
    edge = np.zeros(image.shape[:2], np.uint8); edge[:20, :30] = 1; edge[-5:, -40:] = 1
    assert np.array_equal(telea_inpaint_cropped(image, edge, 1), cv2.inpaint(image, edge, inpaintRadius=1, flags=cv2.INPAINT_TELEA))
    assert np.array_equal(telea_inpaint_cropped(image, np.zeros_like(edge), 1), image)


def test_stem_weight_permutation_matches_patch_order():
    """ops.stem_weight_as_1x1: k = ky*24 + c*8 + kx (kx = 7 and k >= 168 zero) — the order frtm_stem_patches_u8 writes."""
    import torch
    from frtm_vos_b200 import ops
    w = torch.arange(2 * 3 * 7 * 7, dtype=torch.float32).reshape(2, 3, 7, 7)
    p = ops.stem_weight_as_1x1(w)
    assert p.shape == (2, 192, 1, 1)
    for n, c, ky, kx in [(0, 0, 0, 0), (1, 2, 6, 6), (0, 1, 3, 5), (1, 0, 2, 0)]:
        assert p[n, ky * 24 + c * 8 + kx, 0, 0] == w[n, c, ky, kx]
    assert float(p[:, 168:].abs().sum()) == 0.0
    assert float(p[:, 7:168:8].abs().sum()) == 0.0          # the padding element of every (ky, c) chunk
    # a 1x1 conv over explicitly gathered patches equals the 7x7 / stride 2 / pad 3 conv
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 12, 15, generator=g)
    wt = torch.randn(4, 3, 7, 7, generator=g)
    ref = torch.nn.functional.conv2d(x, wt, None, 2, 3)
    xp = torch.nn.functional.pad(x, (3, 3, 3, 3))
    Ho, Wo = ref.shape[-2:]
    patches = torch.zeros(1, 192, Ho, Wo)
    for ky in range(7):
        for c in range(3):
            for kx in range(7):
                patches[0, ky * 24 + c * 8 + kx] = xp[0, c, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2]
    out = torch.nn.functional.conv2d(patches, ops.stem_weight_as_1x1(wt))
    assert (out - ref).abs().max() < 1e-4


def test_operator_image_size_formula():
    """frtm_split_sample_bytes: even tile count of 64 pixels (hi + lo planes) plus 256-pixel stencil chunks."""
    from frtm_vos_b200._lib import lib
    L = lib()
    for c, hw in [(96, 1620), (96, 3600), (64, 117), (96, 28), (112, 136)]:
        ntiles = -(-hw // 128) * 2
        nchunks = -(-hw // 256)
        assert L.split_sample_bytes(c, hw) == ntiles * 2 * c * 64 * 2 + nchunks * 10 * 256 * 4


def test_block_scheduler_never_crosses_a_filter_update_or_an_object_start():
    """Tracker._block_length (host logic): a block of frames goes through the network as one batch only if no live
    object's filter changes inside it — the reference updates a filter when ``frame_num % train_skipping == 0`` after the
    increment in ``apply`` (model/discriminator.py:201-227) — and no object starts inside it (model/tracker.py:165-191)."""
    from types import SimpleNamespace as NS
    from frtm_vos_b200.model.tracker import Tracker
    rng = np.random.RandomState(3)
    for trial in range(200):
        skip = int(rng.choice([1, 3, 8, 16]))
        max_block = int(rng.choice([1, 4, 8, 12]))
        seq_len = int(rng.randint(2, 60))
        starts = sorted(set([0] + [int(x) for x in rng.randint(0, seq_len, size=rng.randint(0, 3))]))
        targets = {i + 1: NS(start_frame=s, discriminator=NS(frame_num=0, train_skipping=skip)) for i, s in enumerate(starts)}
        trk = NS(block_batching=True, max_block=max_block, targets=targets)
        trk._live_at = lambda f, trk=trk: Tracker._live_at(trk, f)
        has_new = lambda j: j < seq_len and j in starts
        first = 0
        while first < seq_len:
            if has_new(first):
                n = 1                                                   # run_sequence processes start frames alone
                assert Tracker._block_length(trk, first, seq_len, has_new) == 1
            else:
                n = Tracker._block_length(trk, first, seq_len, has_new)
            assert 1 <= n <= max(max_block, 1) and first + n <= seq_len
            live = [t for t in targets.values() if t.start_frame < first]
            for k in range(n):
                if k > 0:
                    assert not has_new(first + k)                       # no object appears inside a block
                for t in live:
                    t.discriminator.frame_num += 1                      # Discriminator.apply
                    if t.discriminator.frame_num % skip == 0:           # an update fires on this frame ...
                        assert k == n - 1, (trial, first, k, n)         # ... so it must be the block's last frame
            first += n
    # batching switched off: always single frames
    trk = NS(block_batching=False, max_block=8, targets={})
    trk._live_at = lambda f: []
    assert Tracker._block_length(trk, 3, 10, lambda j: False) == 1


def test_planar_cut_and_inpaint_equals_the_interleaved_formula():
    """cut_and_inpaint(d=1, f=1) works on planar (C,H,W) arrays and interleaves only the hole's box; it must equal the
    reference's formula (model/augmenter.py:297-340 at d = f = 1) evaluated on the full interleaved frame."""
    import cv2
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import cut_and_inpaint
    seq = synth.SyntheticSequence(num_objects=3, num_frames=1, size=GI.MID, seq_id=4)
    im, lb, ids = seq[0]
    masks = [(lb == o).byte() for o in ids] + [torch.zeros_like(lb), torch.ones_like(lb)]
    edge = torch.zeros_like(lb); edge[..., :5, -7:] = 1           # object touching the frame corner
    for mask in masks + [edge]:
        image = im.numpy().transpose((1, 2, 0))
        m = (mask.squeeze() > 0).byte().numpy()[..., None]
        cut_ref = np.concatenate((m * image, m * 255), axis=-1).transpose((2, 0, 1))
        outer = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2)))
        bg_ref = cv2.inpaint(np.ascontiguousarray(image), outer, inpaintRadius=1, flags=cv2.INPAINT_TELEA).transpose((2, 0, 1))
        cut, bg = cut_and_inpaint(im, mask, 1, 1)
        assert cut.dtype == torch.uint8 and bg.dtype == torch.uint8 and cut.is_contiguous() and bg.is_contiguous()
        assert np.array_equal(cut.numpy(), cut_ref) and np.array_equal(bg.numpy(), bg_ref)
    assert torch.equal(im, seq[0][0])                              # the input frame is not modified in place


def test_bench_clock_sampler_degrades_without_a_gpu():
    """bench.ClockSampler (NVML poll thread, nvidia-smi child as the second choice) must not raise where neither can open a
    device — the reference arm and the CPU checks run it on boxes without one — and says so in ``reasons``."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if not torch.cuda.is_available():
        assert out["sm_mhz"] is None
    # the throttle-reason bits are NVML's (nvml.h: nvmlClocksThrottleReason*)
    assert dict((n, b) for b, n in bench.ClockSampler.BITS) == dict(
        hw_slowdown=0x8, hw_thermal_slowdown=0x40, sw_thermal_slowdown=0x20, sw_power_cap=0x4)
