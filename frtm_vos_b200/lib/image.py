"""Image I/O helpers used by ``Tracker.run_dataset`` (``lib/image.py:17-35``): PIL read -> (C,H,W) uint8 tensor,
palette PNG write with the DAVIS colour map."""
import numpy as np
import torch
from PIL import Image


def _davis_palette():
    pal = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, 1)
    # bit-interleaved PASCAL-VOC / DAVIS colour map for the first entries
    for i in range(22):
        r = g = b = 0
        cid = i
        for j in range(8):
            r |= ((cid >> 0) & 1) << (7 - j)
            g |= ((cid >> 1) & 1) << (7 - j)
            b |= ((cid >> 2) & 1) << (7 - j)
            cid >>= 3
        pal[i] = (r, g, b)
    pal[:22][pal[:22] == 192] = 191    # the DAVIS palette file uses 191 where the VOC colour map has 192
    return pal


davis_palette = _davis_palette()


def imread(filename):
    im = np.atleast_3d(np.array(Image.open(filename)))
    return torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1)))


def imwrite_indexed(filename, im, color_palette=None):
    assert len(im.shape) < 4 or im.shape[0] == 1
    pal = davis_palette if color_palette is None else color_palette
    out = Image.fromarray(im.detach().cpu().squeeze().numpy(), "P")
    out.putpalette(pal.ravel())
    out.save(filename)
