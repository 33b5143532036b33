"""-m gpu: device-side rendering of the first-frame augmentation views against the host (OpenCV) rendering that
reproduces the reference bit for bit.  The device bicubic samples at exact coordinates while OpenCV quantises them to
1/32 px with fixed-point coefficients, so images agree to within interpolation noise; masks are identical."""
import numpy as np
import pytest
import torch

import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_device_views_match_host_rendering():
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    seq = synth.SyntheticSequence(num_objects=2, num_frames=1, size=(240, 428), seq_id=5)
    im, lb, _ = seq[0]
    for oid in (1, 2):
        mask = (lb == oid).byte()
        np.random.seed(0); torch.manual_seed(0)
        h_im, h_lb = ImageAugmenter(GI.AUG_PARAMS).augment_first_frame(im, mask)
        np.random.seed(0); torch.manual_seed(0)
        d_im, d_lb = ImageAugmenter(GI.AUG_PARAMS).augment_first_frame(im.to(DEV), mask.to(DEV))
        assert d_im.is_cuda and d_im.shape == h_im.shape and d_im.dtype == torch.uint8
        assert torch.equal(d_lb.cpu(), h_lb)                       # labels come from the same host path
        assert torch.equal(d_im[0].cpu(), im)                      # view 0 is the original frame
        diff = (d_im.cpu().float() - h_im.float()).abs()
        assert diff[1:].mean().item() < 1.0, diff[1:].mean().item()
        assert (diff[1:] > 8).float().mean().item() < 0.02


def test_warp_affine_identity_and_shift():
    import ctypes
    from frtm_vos_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(3)
    src = torch.randint(0, 256, (3, 40, 56), generator=g, dtype=torch.uint8).to(DEV)
    out = torch.empty((3, 40, 56), device=DEV)
    ident = (ctypes.c_double * 6)(1, 0, 0, 0, 1, 0)
    lib().warp_affine(ptr(src), 1, 3, 40, 56, ptr(out), None, 40, 56, ident, 0, 0.0, 255.0, stream())
    assert torch.equal(out.cpu(), src.cpu().float())               # bicubic at integer coordinates is exact
    shift = (ctypes.c_double * 6)(1, 0, 5, 0, 1, -3)               # dst(x,y) = src(x-5, y+3)
    lib().warp_affine(ptr(src), 1, 3, 40, 56, ptr(out), None, 40, 56, shift, 1, 0.0, 255.0, stream())
    ref = torch.zeros(3, 40, 56)
    ref[:, :37, 5:] = src.cpu().float()[:, 3:, :51]
    assert torch.equal(out.cpu(), ref)


def test_device_mask_warp_is_bit_identical_to_cv2():
    import ctypes
    import cv2
    from frtm_vos_b200._lib import lib, ptr, stream
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model import augmenter as A
    seq = synth.SyntheticSequence(num_objects=3, num_frames=1, size=(480, 854), seq_id=2)
    lb = seq[0][1]
    np.random.seed(3)
    for oid in (1, 2, 3):
        mask = (lb == oid).byte()
        bbox = A.mask_center_bbox(mask)
        pool = dict(GI.AUG_PARAMS["fg_aug_params"])
        pool["location"] = A.target_locations(5, (480, 854))
        for spec in A.draw_specs(pool)[:8]:
            T, _ = A.spec_transform(spec, bbox, (480, 854))
            ref = A.warp_affine_host(mask, np.array(T, dtype=np.float32), (480, 854), "nearest")
            M = (ctypes.c_double * 6)(*np.asarray(T, dtype=np.float32)[:2, :].astype(np.float64).ravel())
            out = torch.empty((480, 854), device=DEV, dtype=torch.uint8)
            cnt = torch.zeros(1, device=DEV, dtype=torch.int32)
            lib().warp_mask_nearest(ptr(mask.to(DEV).reshape(480, 854).contiguous()), 480, 854, ptr(out), 480, 854, M, 1,
                                    ptr(cnt), stream())
            assert torch.equal(out.cpu(), ref.reshape(480, 854)), spec
            assert int(cnt) == int((ref == 1).sum())


def test_batched_mask_warp_equals_single_warps():
    """frtm_warp_mask_nearest_batch (all candidate masks of a round in one launch) == n calls of frtm_warp_mask_nearest."""
    import ctypes
    from frtm_vos_b200._lib import lib, ptr, stream
    L = lib()
    g = torch.Generator().manual_seed(4)
    H, W, n = 97, 131, 37                                  # more than one batch of 32
    src = (torch.rand(H, W, generator=g) > 0.6).to(torch.uint8).to(DEV)
    rng = np.random.RandomState(8)
    Ms = []
    for _ in range(n):
        a, s = rng.uniform(-1, 1), rng.uniform(0.5, 1.6)
        Ms.append([s * np.cos(a), -s * np.sin(a), rng.uniform(-20, 20), s * np.sin(a), s * np.cos(a), rng.uniform(-20, 20)])
    Ms = np.asarray(Ms, dtype=np.float32).astype(np.float64)
    out_b = torch.empty(n, H, W, dtype=torch.uint8, device=DEV)
    cnt_b = torch.zeros(n, dtype=torch.int32, device=DEV)
    L.warp_mask_nearest_batch(ptr(src), H, W, ptr(out_b), H, W, n, (ctypes.c_double * (6 * n))(*Ms.ravel()), 1, ptr(cnt_b), stream())
    out_s = torch.empty_like(out_b)
    cnt_s = torch.zeros_like(cnt_b)
    for j in range(n):
        L.warp_mask_nearest(ptr(src), H, W, ptr(out_s[j]), H, W, (ctypes.c_double * 6)(*Ms[j]), 1, cnt_s[j:j + 1].data_ptr(), stream())
    assert torch.equal(out_b, out_s) and torch.equal(cnt_b, cnt_s) and int(cnt_b.sum()) > 0
