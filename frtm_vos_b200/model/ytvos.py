"""The all-frames YouTubeVOS variant of the sequence driver (``ytvos_validation/tracker.py:82-207``,
``evaluate_ytvos_valid_all_frames.py``).

It is the same per-frame path — backbone, projection + correlation, refinement network, merge, frame memory, GN/CG filter
update — driven differently:

  * every object is initialised BEFORE the frame loop, from the frame it first appears on (``:113-131``);
  * a frame returns the objects' raw probabilities ``sigmoid(logits)`` (pixels of objects that start on that frame
    suppressed, ``:133-139``), not merged masks; objects that have not started yet are zero (``:181-184``);
  * the merge that feeds the target-model update sees one row per object of the SEQUENCE (zeros for objects that are not
    live), and with ``pixel_weighting['update_method'] == 'thresh'`` (the variant's configuration) the training label is
    the binarised merged mask (``ytvos_validation/discriminator.py:364-367``);
  * after the last frame the ground truth is re-inserted on every object's first frame and ONE softmax / argmax over
    {background, objects} produces the label maps of all frames (``:99-107``);
  * the refinement network ends in the true-bicubic ``Upsampler`` (``ytvos_validation/seg_network.py:62-74``):
    ``SegNetwork(..., upsampler="bicubic")``.

Everything runs on the kernels of libfrtm_b200; the sequence protocol is this package's (``lib/datasets.py``: items
``(image, labels | [], new_object_ids)``), so ``YouTubeVOSDataset(..., all_annotations=...)`` sequences can be fed directly.
"""
from __future__ import annotations

from time import time

import torch

from .. import ops
from .tracker import Tracker


class YtvosTracker(Tracker):

    def __init__(self, augmenter, feature_extractor, disc_params, refiner, device):
        super().__init__(augmenter, feature_extractor, disc_params, refiner, device)
        pw = disc_params.get("pixel_weighting") or {}
        self.update_method = pw.get("update_method", "thresh")
        if self.update_method not in ("thresh", "raw"):
            raise NotImplementedError("update_method '%s' (implemented: 'thresh', the variant's configuration, and 'raw')"
                                      % self.update_method)

    # ------------------------------------------------------------------------------------------------------------
    def run_sequence(self, sequence, speedrun=False):
        """-> (list of uint8 label maps (1,H,W), frames / second incl. the initialisations)."""
        self.eval()
        self.object_ids = list(sequence.obj_ids)
        ids = self.object_ids
        self.targets = dict()
        self._stack = None
        self._fbuf = None
        self._gn_table = None
        self._lut = torch.tensor([0] + ids, dtype=torch.uint8, device=self.device)
        n_frames = len(sequence)
        items = [sequence[i] for i in range(n_frames)]
        first = {}
        for i, (im, lb, new) in enumerate(items):
            for oid in new:
                first[oid] = i
        missing = [o for o in ids if o not in first]
        if missing:
            raise ValueError("objects %s never start in sequence %s" % (missing, getattr(sequence, "name", "?")))
        t0 = time()
        # ---- every object from its own first frame (objects sharing a frame are initialised together) ----
        for f0 in sorted(set(first.values())):
            group = [o for o in ids if first[o] == f0]
            self.current_frame = f0
            self.initialize(items[f0][0].to(self.device), items[f0][1].to(self.device), group)
        index = {o: k for k, o in enumerate(ids)}                 # row of an object in the per-frame outputs
        size = tuple(items[0][0].shape[-2:])
        out = torch.zeros((n_frames, len(ids), *size), device=self.device)
        # ---- frames ----
        for i in range(n_frames):
            self.current_frame = i
            self._track_frame(items[i][0].to(self.device), index, out[i])
        # ---- ground truth back in, one merge over all frames ----
        for o in ids:
            f0 = first[o]
            out[f0, index[o]] = (items[f0][1].to(self.device)[0] == o).float()
        labels = ops.labels_from_probs(out, self._lut)
        torch.cuda.synchronize()
        T = time() - t0
        return [labels[i].unsqueeze(0) for i in range(n_frames)], n_frames / T

    # ------------------------------------------------------------------------------------------------------------
    def _track_frame(self, image, index, out_row):
        """One frame (``:163-207``): probabilities of the live objects into ``out_row`` (N,H,W), target-model updates."""
        live = [t for t in self.targets.values() if t.start_frame < self.current_frame]
        if not live:
            return
        n = len(live)
        N = len(self.object_ids)
        im_size = image.shape[-2:]
        feats, _, _ = self.feature_extractor.forward_split(image.unsqueeze(0) if image.dim() == 3 else image)
        fmap = feats[live[0].disc_layer]
        h, w = fmap.hi.shape[1:3]
        c = live[0].discriminator.filter.weight.shape[1]
        samples = ops.conv2d_tc(fmap, self._stacked_projection(live), out_f32=False, nchw=True)["nchw"].view(n, c, h, w)
        fidx = torch.tensor([t.index - 1 for t in live], dtype=torch.int32).to(self.device, non_blocking=True)
        scores = ops.corr3x3(samples, self._fbuf, fidx)
        logits = self.refiner.forward_nhwc(scores, feats, im_size).view(n, *im_size)
        fresh = [t for t in self.targets.values() if t.start_frame == self.current_frame]
        suppress = None
        if fresh:
            suppress = torch.stack([t.start_mask.reshape(*im_size) for t in fresh]).amax(dim=0).contiguous()
        probs = ops.sigmoid_suppress(logits, suppress)
        rows = torch.tensor([index[t.object_id] for t in live], dtype=torch.long, device=self.device)
        out_row.index_copy_(0, rows, probs)
        for k, t in enumerate(live):
            t.discriminator.frame_num += 1
            t.discriminator.current_sample = samples[k:k + 1]
        if not self.disc_params["update_filters"] or self.current_frame == 0:
            return
        # the merge of ``update`` (``:133-161``): one row per object of the sequence, zeros for objects that are not live
        src = torch.zeros((N, *im_size), device=self.device)
        src.index_copy_(0, rows, logits)
        mask_bits = 0
        for t in live:
            mask_bits |= 1 << index[t.object_id]
        counts = torch.zeros(N, dtype=torch.int32, device=self.device)
        masks, _, counts = ops.merge_masks(src, mask_bits, suppress, self._lut, False, counts=counts)
        d0 = live[0].discriminator
        tf = d0.pw_params["tf"] if d0.pw_params else 0.0
        hinge = d0.pw_params is not None and d0.pw_params.get("method") == "hinge"
        for t in live:
            r = index[t.object_id]
            y = masks[1 + r].reshape(1, 1, *im_size)
            if self.update_method == "thresh":
                y = ops.threshold(y, 0.5)
            # 'thresh': hinge weights of the binary label; 'raw': unit weights (discriminator.py:364-377)
            pw = ops.pixel_weights(y, tf, True, counts=counts[r:r + 1]) if (hinge and self.update_method == "thresh") else torch.ones_like(y)
            t.discriminator.update(y, gate_count=counts[r:r + 1], pw=pw)
