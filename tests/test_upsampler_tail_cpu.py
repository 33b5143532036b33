"""CPU: the index / weight arithmetic of the fused upsampler tail (csrc/elementwise.cu: fused_axis_window,
upsample_tapsum_kernel) restated in numpy, tile by tile, against the chain it replaces —
conv3x3(bilinear(pyr_up_bicubic(t))) of /root/reference/model/seg_network.py:75-126,138-146 — over sizes the GPU tests do
not sample, together with the window invariants the kernel relies on (its shared-memory reads are not bounds-checked)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

E4 = np.array([-0.10546875, 0.87890625, 0.26171875, -0.03515625], dtype=np.float32)   # kCubicE
TY, TX, NTY, NTX = 16, 64, 16, 42                                                      # FT_TY, FT_TX, FT_NTY, FT_NTX
f32 = np.float32


def bilinear_src(dst, scale, in_size):
    src = f32(scale) * (f32(dst) + f32(0.5)) - f32(0.5)
    if src < 0:
        src = f32(0)
    i0 = min(int(src), in_size - 1)
    i1 = i0 + (1 if i0 < in_size - 1 else 0)
    return i0, i1, f32(src - f32(i0))


def fused_axis_window(d, ok, scale, up_size, lo):
    u0, u1, lam = bilinear_src(d, scale, up_size)
    n0, n1 = (u0 + 1) >> 1, (u1 + 1) >> 1
    w0, w1 = (f32(1) - lam, lam) if ok else (f32(0), f32(0))
    c0 = w0 * (E4[::-1] if (u0 + 1) & 1 else E4)
    c1 = w1 * (E4[::-1] if (u1 + 1) & 1 else E4)
    a = np.zeros(5, dtype=np.float32)
    a[:4] = c0
    assert n1 - n0 in (0, 1)
    a[n1 - n0:n1 - n0 + 4] += c1
    return n0 - 2 - lo, a


def tail_numpy(t, H, W):
    """t (9,h,w) float32 tap maps -> (H,W), one 16x64 tile at a time like the kernel."""
    _, h, w = t.shape
    Hu, Wu = 2 * h, 2 * w
    sy, sx = f32(Hu) / f32(H), f32(Wu) / f32(W)
    out = np.zeros((H, W), dtype=np.float32)
    for Y0 in range(0, H, TY):
        for X0 in range(0, W, TX):
            uy_lo = bilinear_src(max(Y0 - 1, 0), sy, Hu)[0]
            nuy = bilinear_src(min(Y0 + TY, H - 1), sy, Hu)[1] - uy_lo + 1
            ux_lo = bilinear_src(max(X0 - 1, 0), sx, Wu)[0]
            nux = bilinear_src(min(X0 + TX, W - 1), sx, Wu)[1] - ux_lo + 1
            ty_lo = ((uy_lo + 1) >> 1) - 2
            nty = ((uy_lo + nuy) >> 1) + 1 - ty_lo + 1
            tx_lo = ((ux_lo + 1) >> 1) - 2
            ntx = ((ux_lo + nux) >> 1) + 1 - tx_lo + 1
            assert nty <= NTY and ntx <= NTX, (nty, ntx)
            rows = np.clip(np.arange(ty_lo, ty_lo + nty), 0, h - 1)
            cols = np.clip(np.arange(tx_lo, tx_lo + ntx), 0, w - 1)
            T = t[:, rows][:, :, cols]                                   # replicate padding
            R = np.zeros((9, nty, TX), dtype=np.float32)
            for lx in range(TX):
                for dx in range(3):
                    Xs = X0 + lx + dx - 1
                    ok = 0 <= Xs < W and X0 + lx < W
                    bx, ax = fused_axis_window(min(max(Xs, 0), W - 1), ok, sx, Wu, tx_lo)
                    assert 0 <= bx and bx + 3 <= ntx - 1, (bx, ntx)
                    b4 = min(bx + 4, ntx - 1)
                    assert b4 == bx + 4 or ax[4] == 0                    # the clamp only ever acts on a zero weight
                    for dy in range(3):
                        tap = dy * 3 + dx
                        R[tap, :, lx] = ((ax[0] * T[tap, :, bx] + ax[1] * T[tap, :, bx + 1]) +
                                         (ax[2] * T[tap, :, bx + 2] + ax[3] * T[tap, :, bx + 3])) + ax[4] * T[tap, :, b4]
            for r in range(TY):
                if Y0 + r >= H:
                    break
                acc = np.zeros(TX, dtype=np.float32)
                for dy in range(3):
                    Ys = Y0 + r + dy - 1
                    by, ay = fused_axis_window(min(max(Ys, 0), H - 1), 0 <= Ys < H, sy, Hu, ty_lo)
                    assert 0 <= by and by + 3 <= nty - 1, (by, nty)
                    r4 = min(by + 4, nty - 1)
                    assert r4 == by + 4 or ay[4] == 0
                    for dx in range(3):
                        c = R[dy * 3 + dx]
                        acc += ((ay[0] * c[by] + ay[1] * c[by + 1]) + (ay[2] * c[by + 2] + ay[3] * c[by + 3])) + ay[4] * c[r4]
                n = min(TX, W - X0)
                out[Y0 + r, X0:X0 + n] = acc[:n]
    return out


def _supported(h, w, H, W):
    """The library's own host-side check (no device call): does (h,w) -> x2 -> (H,W) fit the kernel's tile windows?"""
    from frtm_vos_b200._lib import lib
    return lib().upsample_tapsum_supported(h, w, H, W) == 1


def test_sizes_outside_the_tile_windows_are_refused():
    assert not _supported(33, 9, 40, 11)       # x2 then a 1.65x reduction: the 16-row tile would need 18 low-resolution rows
    assert not _supported(10, 10, 30, 30)      # enlarging resize after the x2
    assert _supported(240, 428, 480, 854) and _supported(360, 640, 720, 1280)   # the BASELINE frame sizes


@pytest.mark.parametrize("hw,size", [((24, 43), (48, 84)), ((24, 43), (48, 86)), ((37, 50), (73, 99)), ((13, 17), (26, 34)),
                                     ((20, 33), (36, 61)), ((9, 70), (18, 139)), ((11, 40), (20, 75)), ((33, 9), (60, 17)),
                                     ((8, 8), (16, 16)), ((17, 65), (34, 129)), ((20, 90), (36, 165))])
def test_merged_windows_equal_bicubic_then_bilinear_then_shifted_sum(hw, size):
    from oracle import frtm_ref as R
    assert _supported(hw[0], hw[1], size[0], size[1])
    g = torch.Generator().manual_seed(hw[0] * 131 + size[1])
    h, w = hw
    t = torch.randn(1, 9, h, w, generator=g)
    up = F.interpolate(R.pyr_up_bicubic(t), size, mode="bilinear", align_corners=False)       # (1,9,H,W)
    wt = torch.zeros(1, 9, 3, 3)
    for tap in range(9):
        wt[0, tap, tap // 3, tap % 3] = 1.0                                                    # tap map `tap` shifted by (dy,dx)
    ref = F.conv2d(up, wt, None, 1, 1)[0, 0].numpy()
    out = tail_numpy(t[0].numpy(), size[0], size[1])
    assert np.abs(out - ref).max() < 4e-5            # sums of 9 x 25 products of N(0,1) samples in a different order
