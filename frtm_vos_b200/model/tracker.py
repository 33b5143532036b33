"""Sequence driver (drop-in for ``model/tracker.py:16-227``).

``Tracker(augmenter, feature_extractor, disc_params, refiner, device)`` keeps the reference's public surface:
``run_dataset``, ``run_sequence`` (same fps definition: frames / wall-clock of the frame loop including per-object
initialisation, ``:130,159-161``), ``initialize``, ``track``, ``clear``, attributes ``targets``, ``current_masks``,
``current_frame``.  What changes is how a frame executes:

  * one batched pass per stage instead of a Python loop over objects: the 1x1 projections of all live objects are
    stacked along Cout of a single conv, their 3x3 filters live in one contiguous buffer read by one correlation
    launch, and the refinement network sees all objects as one batch;
  * sigmoid + suppression + clamp + softmax-merge + argmax gating + label LUT + the per-object ``> 0.5`` pixel counts
    are one fused kernel;
  * the ``< 10 px`` update gate and the memory replace-index live on the device — no host sync inside the frame loop;
  * ``run_sequence`` pushes up to 8 consecutive frames through backbone / projection / correlation / refinement as ONE
    batch: masks never feed the next frame's forward pass and the filters only change every ``train_skipping`` frames
    (``model/discriminator.py:221``), so the results are those of the frame-by-frame loop (SURVEY.md finding 4).
"""
from __future__ import annotations

from time import time

import numpy as np
import os

import torch
import torch.nn as nn

from .. import ops
from .discriminator import Discriminator
from .seg_network import SegNetwork


class TargetObject:

    def __init__(self, obj_id, disc_params, **kwargs):
        self.object_id = obj_id
        self.discriminator = Discriminator(**disc_params)
        self.disc_layer = disc_params["layer"]
        self.start_frame = None
        self.start_mask = None
        self.index = -1
        for key, val in kwargs.items():
            setattr(self, key, val)

    def initialize(self, ft, mask):
        # reference path (``model/tracker.py:30-31``): ``ft`` is the dict of NCHW maps; the fp32 NHWC working tensor is used
        # when the extractor happened to produce it, else the NCHW map is converted by ``Discriminator.init``
        nhwc = getattr(ft, "nhwc", None)
        x_nhwc = nhwc.get(self.disc_layer) if isinstance(nhwc, dict) else None
        self.discriminator.init(ft[self.disc_layer] if x_nhwc is None else None, mask, x_nhwc=x_nhwc)

    def classify(self, ft):
        return self.discriminator.apply(ft)


class Tracker(nn.Module):

    def __init__(self, augmenter, feature_extractor, disc_params, refiner: SegNetwork, device):
        super().__init__()
        self.augmenter = augmenter
        self.augment = augmenter.augment_first_frame
        self.disc_params = disc_params
        self.feature_extractor = feature_extractor
        self.refiner = refiner
        for m in self.refiner.parameters():
            m.requires_grad_(False)
        self.refiner.eval()
        self.device = device
        self.first_frames = []
        self.current_frame = 0
        self.current_masks = None
        self.num_objects = 0
        self.targets = dict()
        self.object_ids = []
        # host threads preparing the first-frame augmentations of several new objects at once: the OpenCV inpaint (7 ms,
        # GIL released) dominates a job, so one thread per object up to the process's thread budget (bench.py caps
        # torch's thread count at cores / ranks)
        self.augment_workers = max(4, min(12, torch.get_num_threads()))
        self.block_batching = True  # run_sequence batches up to max_block frames between two filter updates
        self.max_block = 8
        # Full blocks (max_block frames, no object starting inside, aligned to the update schedule) are captured ONCE as a
        # CUDA graph and replayed: ~170 launches become one.  The per-slot target-model buffers are pooled across sequences
        # (see _obj_pool below), so the capture made on the first sequence serves every later sequence with the same number
        # of objects and frame size — re-capturing per sequence, as round 1 did, cost more than its seven replays saved.
        # Measured on B200 (profiles/r02_graph_blocks.md): +1.9 % frames/s on config 2, +2.8 % on config 3 at one GPU (the
        # block is GPU-bound there); the point is the host: 20 ms of launch work per sequence disappear, which is what
        # multi-rank runs on shared cores were short of.  FRTM_GRAPH_BLOCKS=0 turns it off; code that instruments the
        # per-kernel path (the oracle-replay harness) sets ``graph_blocks = False`` on its tracker.
        self.graph_blocks = os.environ.get("FRTM_GRAPH_BLOCKS", "1") == "1"
        self._blk_graph = None
        self._stack = None          # cached stacked projection of the live objects
        self._stack_buf = None      # its packed buffers, rewritten in place (stable addresses across sequences)
        self._fbuf = None           # (maxN, c, 3, 3) contiguous filters (each Discriminator.filter.weight is a view)
        self._last_labels = None
        # Per object slot (1, 2, ...): the frame memory and CG state of the target model in that slot, reused by every
        # sequence (``Discriminator.buffer_pool``).  With the filter buffer, the packed projection, the GN pointer table and
        # the label LUT kept in place as well, every device address a track block touches is the same from one sequence to
        # the next, so ONE captured CUDA graph of a full block serves all sequences with the same number of objects.
        self._obj_pool = {}
        self._lut_buf = None
        # First-frame augmentations of the NEXT sequence prepared behind the current one's tracking (``prefetch_init``):
        # {(id(sequence), frame): record}.  Half of an initialisation is host work (OpenCV inpaint of the cut-out object)
        # that depends on the first frame and its ground truth only — both known before the sequence starts.
        self._prefetched = {}
        self.prefetch_next = os.environ.get("FRTM_PREFETCH_INIT", "1") == "1"
        # ``Memory.labels`` / ``Memory.pixel_weights`` (cap,1,H,W) mirror the reference's buffers (model/memory.py:20-21),
        # but nothing on this path reads them after the insert: the optimiser works on the stencil / U^T w^2 y built from
        # them at insert time.  The block inserts therefore skip the two full-resolution copies (84 % of an insert's bytes)
        # unless somebody wants the mirrors kept up to date: ``store_fullres_memory = True`` (or
        # FRTM_STORE_FULLRES_MEMORY=1).  The first-frame samples and the frame-by-frame path always store them.
        self.store_fullres_memory = os.environ.get("FRTM_STORE_FULLRES_MEMORY", "0") == "1"
        self.graph_captures = 0     # how many times a block graph was captured (tests / diagnostics)

    def _drain_prefetch(self):
        """Wait until the worker threads have issued everything they prepare for coming sequences (host side only)."""
        for rec in list(self._prefetched.values()):
            for f in [rec["features"]]:
                try:
                    f.result()
                except Exception:                          # noqa: BLE001 - surfaces in initialize, where it is consumed
                    pass

    def clear(self):
        self.first_frames = []
        self.current_frame = 0
        self.current_masks = None
        self.num_objects = 0
        self._stack = None
        self._drain_prefetch()      # no cudaFree under the worker threads' launches
        torch.cuda.empty_cache()

    def release_buffers(self):
        """Drop everything cached across sequences: the staging sets of the joint optimisation with their captured CUDA
        graphs (``model/optimizer.py`` keeps a small LRU of them, one per feature resolution and stream) and the
        per-sequence tables.  Not needed between sequences; for long-lived processes that switch models or resolutions."""
        from .optimizer import release_init_stages
        self._drain_prefetch()
        self._prefetched = {}
        self.targets = dict()
        self._fbuf = None
        self._gn_table = None
        self._blk_graph = None
        self._stack = self._stack_buf = None
        self._gn_table_buf = None
        self._obj_pool = {}
        release_init_stages()
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------------------------
    def run_dataset(self, dataset, out_path, speedrun=False, restart=None):
        """Runs every sequence of ``dataset`` and writes indexed PNG label maps under ``out_path`` (``:68-101``)."""
        from ..lib.image import imwrite_indexed
        out_path.mkdir(exist_ok=True, parents=True)
        fps_sum, fps_n = 0.0, 0
        print("Evaluating", dataset.name)
        restarted = False
        n_seq = len(dataset)
        for k in range(n_seq):
            sequence = dataset[k]
            if restart is not None and not restarted:
                if sequence.name != restart:
                    continue
                restarted = True
            sequence.preload(self.device)
            # I/O overlap (lib/datasets.py here): JPEG decode + H2D of the next sequence run behind this one's tracking.
            # Sequences of the reference's own dataset classes have no preload_async and are simply preloaded in turn.
            if k + 1 < n_seq:
                nxt = dataset[k + 1]
                if hasattr(nxt, "preload_async"):
                    nxt.preload_async(self.device)
            self.clear()
            if k + 1 < n_seq and getattr(self, "prefetch_next", False):
                outputs, seq_fps = self.run_sequence(sequence, speedrun, next_sequence=dataset[k + 1])
            else:
                outputs, seq_fps = self.run_sequence(sequence, speedrun)
            if not np.isnan(seq_fps):
                fps_sum, fps_n = fps_sum + seq_fps, fps_n + 1
            dst = out_path / sequence.name
            dst.mkdir(exist_ok=True)
            for lb, f in zip(outputs, sequence.frame_names):
                imwrite_indexed(dst / (f + ".png"), lb)
            if hasattr(sequence, "release"):
                sequence.release()
        print("Average frame rate: %.2f fps" % (fps_sum / max(fps_n, 1)))

    def run_sequence(self, sequence, speedrun=False, next_sequence=None, host_labels=None, sync=True):
        """``next_sequence`` (optional, not in the reference signature): the sequence that will be run after this one; the
        host part of its first-frame initialisation (augmentation) is prepared in worker threads while this sequence is
        tracked.  Results are identical with and without it.  ``host_labels`` (optional): a pinned uint8 tensor
        (frames, H, W); every frame's label map is also copied into it on a copy stream as soon as its block is done, so the
        device->host transfer of the results runs behind the tracking instead of after it (complete on return).
        ``sync=False``: do not wait for the device at the end — the returned label maps (and ``host_labels``) are complete once
        ``self.sequence_done`` (a CUDA event) has been reached, the frame rate is not measured (nan) — so that a driver can
        issue the next sequence at once: its initialisation then overlaps this sequence's last block."""
        self.eval()
        self.object_ids = sequence.obj_ids
        self.current_frame = 0
        self.targets = dict()
        self._stack = None
        lut = [0] + list(sequence.obj_ids)
        if self._lut_buf is None or self._lut_buf.numel() < len(lut):
            self._lut_buf = torch.zeros(max(len(lut), 16), dtype=torch.uint8, device=self.device)
        ops.fill_u8(self._lut_buf, lut)                 # in place: the LUT's address is part of a captured block graph
        self._lut = self._lut_buf[:len(lut)]
        N = 0
        if speedrun:
            image, labels, obj_ids = sequence[0]
            image = image.to(self.device)
            labels = labels.to(self.device)
            self.initialize(image, labels, sequence.obj_ids)
            self.track(image)
            torch.cuda.synchronize()
            self.targets = dict()
            self._stack = None

        outputs = []
        single = len(sequence.obj_ids) == 1
        new_cache = {}

        def item(i):
            if i not in new_cache:
                new_cache[i] = sequence[i]
            return new_cache[i]

        # Frames that live on the host are uploaded one block AHEAD on a copy stream: the H2D of block b+1 runs behind the
        # kernels of block b (and the first block's behind the initialisation) instead of in front of its own block.
        ahead = {}
        main = torch.cuda.current_stream()

        def upload_ahead(j0, j1):
            for j in range(j0, min(j1, n_frames)):
                im = item(j)[0]
                if j in ahead or not torch.is_tensor(im) or im.is_cuda:
                    continue
                if getattr(self, "_h2d_stream", None) is None:
                    self._h2d_stream = torch.cuda.Stream(device=self.device)
                with torch.cuda.stream(self._h2d_stream):
                    d = im.to(self.device, non_blocking=True)
                    ahead[j] = (d, self._h2d_stream.record_event())

        def to_host(j, lab):
            if host_labels is None:
                return
            if getattr(self, "_d2h_stream", None) is None:
                self._d2h_stream = torch.cuda.Stream(device=self.device)
            self._d2h_stream.wait_stream(main)
            with torch.cuda.stream(self._d2h_stream):
                host_labels[j].copy_(lab.reshape(host_labels.shape[-2:]), non_blocking=True)
            lab.record_stream(self._d2h_stream)

        def frame(j):
            if j in ahead:
                d, ev = ahead.pop(j)
                main.wait_event(ev)
                d.record_stream(main)
                return d
            return item(j)[0].to(self.device, non_blocking=True)

        t0 = time()
        i = 0
        n_frames = len(sequence)
        while i < n_frames:
            image, labels, new_objects = item(i)
            had_targets = len(self.targets) > 0
            if len(new_objects) > 0 or not had_targets:
                upload_ahead(i + 1, i + 1 + self.max_block)
                image = frame(i)
                if len(new_objects) > 0:
                    labels = labels.to(self.device)
                    pf = self._prefetched.pop((id(sequence), i), None)
                    if pf is None:
                        self.initialize(image, labels, new_objects)        # the reference's call
                    else:
                        self.initialize(image, labels, new_objects, prefetched=pf)
                if had_targets:
                    self.track(image)
                    labels = self._last_labels.unsqueeze(0) if single else self._last_labels
                if isinstance(labels, list) and len(labels) == 0:
                    labels = image.new_zeros(1, *image.shape[-2:])
                outputs.append(labels)
                to_host(i, labels)
                self.current_frame += 1
                new_cache.pop(i, None)
                i += 1
                N += 1
                continue
            nb = self._block_length(i, n_frames, lambda j: j < n_frames and len(item(j)[2]) > 0)
            images = [frame(j) for j in range(i, i + nb)]
            if next_sequence is not None and self.prefetch_next:
                self.prefetch_init(next_sequence)         # once: behind the first track block of this sequence
                next_sequence = None
            for f, lab in enumerate(self._track_block(images)):
                outputs.append(lab.unsqueeze(0) if single else lab)
                to_host(i + f, lab)
            upload_ahead(i + nb, i + nb + self.max_block)
            for j in range(i, i + nb):
                new_cache.pop(j, None)
            self.current_frame += nb
            i += nb
            N += nb
        if getattr(self, "_d2h_stream", None) is not None and host_labels is not None:
            main.wait_stream(self._d2h_stream)
        self.sequence_done = main.record_event()
        if not sync:
            return outputs, float("nan")
        torch.cuda.synchronize()
        T = time() - t0
        return outputs, N / T

    # ------------------------------------------------------------------------------------------------------------
    def _bind_filter(self, target):
        """Move the object's 3x3 filter into the shared contiguous buffer (one correlation launch for all objects)."""
        d = target.discriminator
        c = d.filter.weight.shape[1]
        need = target.index
        if self._fbuf is None or self._fbuf.shape[0] < need:
            cap = max(need, len(self.object_ids), 4)
            new = torch.zeros(cap, c, 3, 3, device=d.filter.weight.device)
            if self._fbuf is not None:
                new[:self._fbuf.shape[0]] = self._fbuf
                for t in self.targets.values():
                    if t is not target:
                        t.discriminator.filter.weight.data = new[t.index - 1:t.index]
            self._fbuf = new
        self._fbuf[need - 1:need] = d.filter.weight.data
        d.filter.weight.data = self._fbuf[need - 1:need]
        if d.update_optimizer is not None:
            d.update_optimizer.x[0] = d.filter.weight

    def _augment_uses_rng(self):
        if getattr(self, "_augment_takes_rng", None) is None:
            import inspect
            try:
                self._augment_takes_rng = "rng" in inspect.signature(self.augment).parameters
            except (TypeError, ValueError):
                self._augment_takes_rng = False
        return self._augment_takes_rng

    def _aug_pool(self):
        if getattr(self, "_pool", None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.augment_workers, thread_name_prefix="frtm-aug")
        return self._pool

    def prefetch_init(self, sequence, frame=0):
        """Start the first-frame augmentations of ``sequence`` (every object that starts in ``frame``) in the worker threads,
        on side streams of their own; ``run_sequence(sequence)`` picks the views up in ``initialize``.  The views depend on
        the frame, its ground truth and a per-object ``RandomState(0)`` only (``model/tracker.py:178-183``), so they are
        the ones ``initialize`` would compute itself."""
        key = (id(sequence), frame)
        if key in self._prefetched or not self._augment_uses_rng():
            return
        image, labels, new_objects = sequence[frame]
        if len(new_objects) == 0:
            return
        n_new = len(new_objects)
        main = torch.cuda.current_stream()
        if getattr(self, "_pf_streams", None) is None or len(self._pf_streams) < n_new:
            self._pf_streams = [torch.cuda.Stream(device=self.device) for _ in range(n_new)]
        up = self._pf_streams[0]
        up.wait_stream(main)
        with torch.cuda.stream(up):                   # the first frame may still be on the host (pinned): upload it here
            image_d = image.to(self.device, non_blocking=True)
            labels_d = labels.to(self.device, non_blocking=True)
            uploaded = up.record_event()

        def job(k):
            side = self._pf_streams[k]
            with torch.cuda.stream(side):
                side.wait_event(uploaded)
                image_d.record_stream(side); labels_d.record_stream(side)
                mask = (labels_d == new_objects[k]).byte()
                im, msk = self.augment(image_d, mask, rng=np.random.RandomState(0))
                im, msk = im.to(self.device), msk.to(self.device)
                return im, msk, side.record_event()

        pool = self._aug_pool()
        futures = [pool.submit(job, k) for k in range(n_new)]
        layer = self.disc_params["layer"]

        def features():
            # the views of all objects through the backbone as ONE batch (the pass does not depend on the target models,
            # which ``initialize`` draws and fits at the reference's place in the random stream)
            res = [f.result() for f in futures]
            side = self._pf_streams[0]
            with torch.cuda.stream(side):
                for im, msk, ev in res:
                    side.wait_event(ev)
                    im.record_stream(side)
                _, f32, _ = self.feature_extractor.forward_split(torch.cat([r[0] for r in res]), (), (layer,), upto=layer)
                x, done = f32[layer], side.record_event()
            nv = res[0][0].shape[0]
            return [(res[k][0], res[k][1], done, x[k * nv:(k + 1) * nv]) for k in range(n_new)]

        rec = dict(sequence=sequence, objects=tuple(new_objects), keep=(image_d, labels_d), uploaded=uploaded, futures=futures,
                   features=pool.submit(features))
        self._prefetched[key] = rec
        while len(self._prefetched) > 2:              # an entry whose sequence never ran
            self._prefetched.pop(next(iter(self._prefetched)))

    def _reseed0(self):
        # torch.random.manual_seed(0) of the reference (``:178``), without its detour through the lazy-init queues of back
        # ends that are not in use (each queued call formats a stack trace: 0.3 ms per object)
        torch.default_generator.manual_seed(0)
        if torch.cuda.is_initialized():
            torch.cuda.manual_seed_all(0)

    def initialize(self, image, labels, new_objects, prefetched=None):
        """Create and fit a target model per new object (``:165-191``).

        Augmentation is host work (OpenCV inpaint, spec drawing) plus a few device kernels and one small device->host read
        per object.  It runs in worker threads, each on its own CUDA side stream and with its own
        ``numpy.random.RandomState(0)`` (the reference reseeds the global generators to 0 before every object,
        ``:178-180``, so the draws are identical), while the main thread fits the objects on the main stream in order."""
        self.current_masks = torch.zeros((len(self.targets) + len(new_objects) + 1, *image.shape[-2:]), device=self.device)
        main = torch.cuda.current_stream()
        n_new = len(new_objects)
        if getattr(self, "_aug_streams", None) is None or len(self._aug_streams) < n_new:
            self._aug_streams = [torch.cuda.Stream(device=image.device) for _ in range(max(n_new, 1))]
        matches = prefetched is not None and prefetched["objects"] == tuple(new_objects)
        if matches:
            # Views and features were prepared on streams of their own: the joint fits below depend on nothing the main
            # stream still has queued, so they start at once — while the PREVIOUS sequence's last block is still running
            # when the caller chains sequences with ``run_sequence(..., sync=False)`` — and only the second half of an
            # initialisation (the pooled frame memory of the object slot) waits for the main stream.
            labels = prefetched["keep"][1]           # the same ground truth, uploaded by prefetch_init on its own stream
            for st in self._aug_streams[:n_new]:
                st.wait_event(prefetched["uploaded"])
        else:
            for st in self._aug_streams[:n_new]:
                st.wait_stream(main)                 # image / labels uploads are visible to the side streams
        # targets are constructed in order on the main thread: each constructor draws its initial weights from the global
        # torch generator exactly where the reference does (before the reseed of that object)
        targets = []
        for k, obj_id in enumerate(new_objects):
            with torch.cuda.stream(self._aug_streams[k]):
                mask = (labels == obj_id).byte()
                # built under a side stream: the (pageable, hence synchronous) upload of the freshly drawn project/filter
                # weights then waits for that stream only
                target = TargetObject(obj_id=obj_id, index=len(self.targets) + 1, disc_params=self.disc_params,
                                      start_frame=self.current_frame, start_mask=mask)
            target.discriminator.buffer_pool = self._obj_pool.setdefault(target.index, {})
            self.targets[obj_id] = target
            targets.append(target)
            self._reseed0()
            np.random.seed(0)

        def augment_job(k):
            side = self._aug_streams[k]
            with torch.cuda.stream(side):
                if self._augment_takes_rng:
                    im, msk = self.augment(image, targets[k].start_mask, rng=np.random.RandomState(0))
                else:
                    im, msk = self.augment(image, targets[k].start_mask)
                im, msk = im.to(self.device), msk.to(self.device)
                return im, msk, side.record_event()

        workers = min(n_new, self.augment_workers) if self._augment_uses_rng() else 1
        if matches:
            results = iter(prefetched["features"].result())                # prepared behind the previous sequence
        elif workers > 1:
            pool = self._aug_pool()
            futures = [pool.submit(augment_job, k) for k in range(n_new)]
            results = (f.result() for f in futures)
        else:
            def sequential():
                for k in range(n_new):
                    np.random.seed(0)                # foreign augmenters use the global generator, like the reference
                    yield augment_job(k)
            results = sequential()
        # Object k is fitted on its own stream k (the one its augmentation ran on): the fits are chains of small kernels
        # (~650 per object, replayed as one CUDA graph each) that leave most of the GPU idle, so the objects of a frame
        # overlap on the device instead of queueing behind each other.
        for k, (target, (im, msk, ready, *feat)) in enumerate(zip(targets, results)):
            side = self._aug_streams[k]
            with torch.cuda.stream(side):
                side.wait_event(ready)               # a no-op unless the views come from prefetch_init's streams
                im.record_stream(side); msk.record_stream(side)
                if feat:                             # prefetched: the backbone pass is done as well
                    x = feat[0]
                    x.record_stream(side)
                    target.discriminator.init_joint(None, msk, x_nhwc=x)
                    side.wait_stream(main)           # the slot's pooled memory may still serve the previous sequence
                    target.discriminator.init_memory()
                else:
                    _, f32, _ = self.feature_extractor.forward_split(im, (), (target.disc_layer,), upto=target.disc_layer)
                    x = f32[target.disc_layer]
                    target.discriminator.init(None, msk, x_nhwc=x)
        for k, target in enumerate(targets):
            main.wait_stream(self._aug_streams[k])
            d = target.discriminator
            for t in (target.start_mask, d.project.weight.data, d.filter.weight.data, d.memory.samples, d.memory.labels,
                      d.memory.pixel_weights, d.memory.weights, d.memory.stencil, d.memory.uty, d.memory.state,
                      d.update_optimizer.cg_state):
                t.record_stream(main)
            self._bind_filter(target)
            self.current_masks[target.index] = target.start_mask
        self._stack = None
        self._gn_table = None
        return self.current_masks

    def _live(self):
        return [t for t in self.targets.values() if t.start_frame < self.current_frame]

    def _stacked_projection(self, live):
        key = tuple(t.object_id for t in live)
        if self._stack is None or self._stack[0] != key:
            W = torch.cat([t.discriminator.project.weight.detach() for t in live], dim=0)   # (N*c, C, 1, 1)
            self._stack_buf = ops.pack_conv_tc_1x1_device(W, out=self._stack_buf)
            self._stack = (key, self._stack_buf)
        return self._stack[1]

    def track(self, image):
        """One frame (reference signature, ``:193-227``) = a block of one."""
        self._track_block([image])
        return self.current_masks

    def _block_length(self, first, sequence_len, has_new):
        """How many consecutive frames starting at ``first`` can go through the network as ONE batch with the
        reference's semantics: masks never feed the next frame's forward pass, only the filters do, and those change
        only when some live object reaches a multiple of ``train_skipping`` (SURVEY.md finding 4).  Frames on which
        objects appear are processed alone."""
        if has_new(first) or not self.block_batching:
            return 1
        live = self._live_at(first)
        n = 0
        while first + n < sequence_len and n < self.max_block:
            if n > 0 and has_new(first + n):
                break
            n += 1
            if any((t.discriminator.frame_num + n) % t.discriminator.train_skipping == 0 for t in live):
                break
        return max(n, 1)

    def _frame_index(self, nF, n, dev):
        """Object index of every (frame, object) row of a block, one cached tensor per block shape (its address is part of
        the captured graph of that block length)."""
        if getattr(self, "_fidx", None) is None:
            self._fidx = {}
        t = self._fidx.get((nF, n))
        if t is None or t.device != torch.device(dev):
            t = self._fidx[(nF, n)] = torch.arange(n, dtype=torch.int32, device=dev).repeat(nF).contiguous()
        return t

    def _live_at(self, frame):
        return [t for t in self.targets.values() if t.start_frame < frame]

    def _track_block(self, images):
        """Tracks ``len(images)`` consecutive frames in one batched pass; returns the list of uint8 label maps."""
        nF = len(images)
        if self.graph_blocks and 1 < nF <= self.max_block and self.disc_params["update_filters"] \
                and not any(t.start_frame == self.current_frame for t in self.targets.values()):
            out = self._track_block_graphed(images)
            if out is not None:
                return out
        return self._track_block_eager(images)

    def _track_block_graphed(self, images):
        """Replay (or capture, the first time this sequence sees a full block) the CUDA graph of one full block."""
        live = self._live()
        nF, n = len(images), len(live)
        d0 = live[0].discriminator
        if d0.pw_params is None or d0.pw_params["method"] != "hinge":
            return None
        if any(t.discriminator.frame_num % t.discriminator.train_skipping != 0 for t in live):
            return None                                    # not aligned to the update schedule: eager
        # everything the eager path creates lazily and caches on the tracker exists BEFORE the key is taken / the capture
        # starts, so no long-lived tensor comes out of the graph's private memory pool
        dev = images[0].device
        if getattr(self.refiner, "_packed", 0) is None:
            self.refiner._pack()                          # host-side weight packing uploads pageable tensors
        fidx = self._frame_index(nF, n, dev)
        if getattr(self, "_counts", None) is None or self._counts.numel() < n:
            self._counts = torch.zeros(max(n, 8), dtype=torch.int32, device=dev)
            self._gn_table = None
        if getattr(self, "_counts_blk", None) is None or self._counts_blk.shape[0] < nF or self._counts_blk.shape[1] != n:
            self._counts_blk = torch.zeros((self.max_block, n), dtype=torch.int32, device=dev)
        self._ensure_gn_table(live, list(range(n)))
        self._ins_table = None                            # rebuilt (in place when the object count is unchanged) inside the block
        pc = self._stacked_projection(live)
        # every device address the block's launches carry: the graph is replayed only while all of them are unchanged
        addr = []
        for t in live:
            d, m = t.discriminator, t.discriminator.memory
            addr += [t.index, m.samples.data_ptr(), m.labels.data_ptr(), m.pixel_weights.data_ptr(), m.weights.data_ptr(),
                     m.stencil.data_ptr(), m.uty.data_ptr(), m.split.data_ptr(), m.state.data_ptr(),
                     d.update_optimizer.cg_state.data_ptr(), d.filter.weight.data_ptr()]
        key = (nF, n, tuple(images[0].shape[-2:]), tuple(addr), self._fbuf.data_ptr(), pc.wt.data_ptr(), pc.oscale.data_ptr(),
               self._lut.data_ptr(), len(self.object_ids) == 1, tuple(int(v) for v in d0.update_iters), self._counts.data_ptr(),
               self._counts_blk.data_ptr(), fidx.data_ptr(), self._gn_table[1].data_ptr(), self._gn_table[2].data_ptr())
        if not isinstance(self._blk_graph, dict):
            self._blk_graph = {}                           # one graph per block length (full blocks and the aligned tail)
        g = self._blk_graph.get(nF)
        frames = [im if im.dim() == 3 else im[0] for im in images]
        main = torch.cuda.current_stream()
        if g is None or g["key"] != key:
            self._blk_graph.pop(nF, None)
            static_in = torch.stack(frames)
            side = torch.cuda.Stream(device=dev)
            graph = torch.cuda.CUDAGraph()
            side.wait_stream(main)
            frame_nums = [t.discriminator.frame_num for t in live]
            from .._lib import lib
            launches0 = lib().launch_count()
            try:
                with torch.cuda.stream(side):
                    graph.capture_begin(capture_error_mode="relaxed")
                    try:
                        self._track_block_eager([static_in[f] for f in range(nF)], stacked=static_in)
                    finally:
                        graph.capture_end()
            except Exception as e:                                           # noqa: BLE001 - any capture problem -> eager
                import warnings
                if os.environ.get("FRTM_GRAPH_DEBUG"):
                    import traceback
                    traceback.print_exc()
                warnings.warn("frtm_vos_b200: CUDA-graph capture of the track block failed (%s); running eagerly" % (e,))
                self.graph_blocks = False                  # do not try again on this tracker
                for t, fn in zip(live, frame_nums):
                    t.discriminator.frame_num = fn
                self._gn_table = None
                return None
            main.wait_stream(side)
            # capture records, it does not execute: undo the host-side bookkeeping of the capture pass, then replay
            for t, fn in zip(live, frame_nums):
                t.discriminator.frame_num = fn
            g = dict(key=key, graph=graph, static_in=static_in, labels=self._blk_labels_all, masks=self.current_masks,
                     samples=[t.discriminator.current_sample for t in live], keep=(side,),
                     kernels=int(lib().launch_count() - launches0))
            self._blk_graph[nF] = g
            self.graph_captures += 1
        else:
            torch.stack(frames, out=g["static_in"])
        g["graph"].replay()
        from .._lib import lib
        lib().count_launches(g["kernels"])                # a replay runs the captured kernels without passing the counters
        labels = g["labels"].clone()                      # the static buffers are overwritten by the next replay
        for t, cs in zip(live, g["samples"]):
            t.discriminator.frame_num += nF
            t.discriminator.current_sample = cs
        self.current_masks = g["masks"]
        self._last_labels = labels[nF - 1]
        return [labels[f] for f in range(nF)]

    def _track_block_eager(self, images, stacked=None):
        nF = len(images)
        im_size = images[0].shape[-2:]
        if stacked is not None:
            batch = stacked
        else:
            batch = torch.stack([im if im.dim() == 3 else im[0] for im in images]) if nF > 1 else \
                (images[0] if images[0].dim() == 4 else images[0].unsqueeze(0))
        live = self._live()
        n = len(live)
        if n == 0:
            # Nothing is being tracked yet (``track`` called on the frame the objects were initialised on, as the speedrun
            # warm-up of ``run_sequence`` does, ``model/tracker.py:120-124``): the reference skips its per-object loops
            # and only merges the start masks (``:208-221``).  The backbone pass feeds nothing, so it is skipped.
            fresh = [t for t in self.targets.values() if t.start_frame == self.current_frame]
            assert nF == 1, "a block without live objects is a single frame"
            if not fresh:
                self.current_masks = torch.zeros((1, *im_size), device=batch.device)
                self._last_labels = torch.zeros(im_size, dtype=torch.uint8, device=batch.device)
                return [self._last_labels]
            src = torch.stack([t.start_mask.reshape(*im_size).float() for t in fresh])
            masks, labels, _ = ops.merge_masks(src, 0, None, self._lut, len(self.object_ids) == 1)
            self.current_masks = masks
            self._last_labels = labels
            return [labels]
        feats, _, _ = self.feature_extractor.forward_split(batch)
        layer = live[0].disc_layer
        c = live[0].discriminator.filter.weight.shape[1]
        fmap = feats[layer]
        h, w = fmap.hi.shape[1:3]
        dev = fmap.hi.device

        # classify: one conv for all projections of all frames, one correlation for all filters
        samples = ops.conv2d_tc(fmap, self._stacked_projection(live), out_f32=False, nchw=True)["nchw"].view(nF * n, c, h, w)
        scores = ops.corr3x3(samples, self._fbuf, self._frame_index(nF, n, dev))
        logits = self.refiner.forward_nhwc(scores, feats, im_size).view(nF, n, *im_size)

        fresh = [t for t in self.targets.values() if t.start_frame == self.current_frame]
        assert not fresh or nF == 1
        total = n + len(fresh)
        if getattr(self, "_counts", None) is None or self._counts.numel() < total:
            self._counts = torch.zeros(max(total, 8), dtype=torch.int32, device=dev)   # stable address: it is
            self._gn_table = None                                                      # referenced by the GN table
        d0 = live[0].discriminator
        hinge = d0.pw_params is not None and d0.pw_params["method"] == "hinge"
        out_labels = []
        if nF > 1 and not fresh and self.disc_params["update_filters"] and hinge:
            # A block of frames without new objects: the merges, pixel weights and stencils of all its frames are
            # independent (masks never feed the next frame's forward pass) -> one launch each for the whole block.  Only the
            # memory inserts stay per frame (the slot policy is sequential per object).
            if getattr(self, "_counts_blk", None) is None or self._counts_blk.shape[0] < nF or self._counts_blk.shape[1] != n:
                self._counts_blk = torch.zeros((self.max_block, n), dtype=torch.int32, device=dev)
            cblk = self._counts_blk[:nF]
            masks_all, labels_all = ops.merge_masks_frames(logits.contiguous(), (1 << n) - 1, self._lut, len(self.object_ids) == 1, cblk)
            ys_all = masks_all[:, 1:1 + n].reshape(nF * n, 1, *im_size)
            pw_all = ops.pixel_weights(ys_all, d0.pw_params["tf"], True, counts=cblk.reshape(-1))
            st_all, uty_all = ops.build_stencil(pw_all, ys_all, (h, w))
            # the memory inserts of the whole block: the slot policy of every object runs over its nF frames in one
            # launch (sequential per object, as ``Discriminator.update`` -> ``Memory.update`` would apply it frame by frame),
            # then one launch copies all samples (48 launches per 8-frame block of 3 objects become 2; +1 % frames/s)
            self._insert_block(live, samples, ys_all, pw_all, st_all, uty_all, cblk, nF)
            for f in range(nF):
                out_labels.append(labels_all[f])
            for k, t in enumerate(live):
                t.discriminator.frame_num += nF
                t.discriminator.current_sample = samples[(nF - 1) * n + k:(nF - 1) * n + k + 1]
            self._counts[:n].copy_(cblk[nF - 1])          # the GN table gates on this (stable) buffer
            due = [k for k, t in enumerate(live) if t.discriminator.frame_num % t.discriminator.train_skipping == 0]
            if due:
                self._batched_gn_update(live, due)
            self.current_masks = masks_all[nF - 1]
            self._last_labels = labels_all[nF - 1]
            self._blk_labels_all = labels_all
            return out_labels
        for f in range(nF):
            # merge (objects initialised on this frame take part with their start masks and suppress the others)
            if fresh:
                src = torch.cat([logits[f]] + [t.start_mask.reshape(1, *im_size).float() for t in fresh], dim=0)
                suppress = torch.stack([t.start_mask.reshape(*im_size) for t in fresh]).amax(dim=0).contiguous()
            else:
                src, suppress = logits[f], None
            masks, labels, counts = ops.merge_masks(src, (1 << n) - 1, suppress, self._lut, len(self.object_ids) == 1,
                                                    counts=self._counts[:total])
            out_labels.append(labels)
            for k, t in enumerate(live):
                t.discriminator.frame_num += 1
                t.discriminator.current_sample = samples[f * n + k:f * n + k + 1]
            # learn: memory insert every frame, filter update when due (always the last frame of a block)
            if self.disc_params["update_filters"]:
                ys = masks[1:1 + n].reshape(n, 1, *im_size)
                pw = ops.pixel_weights(ys, d0.pw_params["tf"], True, counts=counts) if hinge else torch.ones_like(ys)
                stencil, uty = ops.build_stencil(pw, ys, (h, w))
                for k, t in enumerate(live):
                    t.discriminator.update(ys[k:k + 1], gate_count=counts[k:k + 1], pw=pw[k:k + 1],
                                           stencil=stencil[k:k + 1], uty=uty[k:k + 1], run_optimizer=False)
                due = [k for k, t in enumerate(live) if t.discriminator.frame_num % t.discriminator.train_skipping == 0]
                if due:
                    assert f == nF - 1, "a filter update inside a block would change the following frames"
                    self._batched_gn_update(live, due)
            self.current_masks = masks
            self._last_labels = labels
        return out_labels

    def _insert_block(self, live, samples, ys_all, pw_all, st_all, uty_all, cblk, nF):
        """``Memory.update`` of every live object for the nF frames of a block (``model/memory.py:59-92``) in two launches."""
        import ctypes
        from .._lib import lib, ptr, stream
        n = len(live)
        mems = [t.discriminator.memory for t in live]
        m0, d0 = mems[0], live[0].discriminator
        full = bool(self.store_fullres_memory)
        key = (full,) + tuple((m.samples.data_ptr(), m.labels.data_ptr(), m.pixel_weights.data_ptr(), m.state.data_ptr()) for m in mems)
        tab = getattr(self, "_ins_table", None)
        if tab is None or tab[0] != key:
            rows = [[m.samples.data_ptr() for m in mems], [m.labels.data_ptr() if full else 0 for m in mems],
                    [m.pixel_weights.data_ptr() if full else 0 for m in mems], [m.stencil.data_ptr() for m in mems],
                    [m.uty.data_ptr() for m in mems], [m.split.data_ptr() if m._split_ok else 0 for m in mems],
                    [m.weights.data_ptr() for m in mems], [m.state.data_ptr() for m in mems]]
            flat = [v for r in rows for v in r]
            old = getattr(self, "_ins_table_buf", None)
            if old is not None and old[0].numel() == len(flat) and old[0].device == samples.device:
                table, slots = old                       # same object count: the addresses stay (captured block graphs)
            else:
                table = torch.empty(len(flat), dtype=torch.int64, device=samples.device)
                slots = torch.empty(self.max_block * n, dtype=torch.int32, device=samples.device)
                self._ins_table_buf = (table, slots)
            lib().fill_i64(ptr(table), (ctypes.c_int64 * len(flat))(*flat), len(flat), stream())   # no synchronising H2D
            self._ins_table = tab = (key, table, slots)
        _, table, slots = tab
        hw = m0.uty.shape[-1] * m0.uty.shape[-2]
        HW = m0.labels.shape[-1] * m0.labels.shape[-2]
        lib().memory_insert_block(ptr(table), n, nF, m0.capacity, float(m0.learning_rates), ptr(cblk), int(d0.min_px),
                                  ptr(samples), samples[0].numel(), ptr(ys_all), ptr(pw_all), HW, ptr(st_all), ptr(uty_all), hw,
                                  1 if m0._split_ok else 0, 1 if self.store_fullres_memory else 0, ptr(slots), stream())

    def _batched_gn_update(self, live, due):
        """One set of launches for the filter updates of all objects that are due on this frame (grid.y = object)."""
        import ctypes
        from .._lib import lib, ptr, stream
        self._ensure_gn_table(live, due)
        _, table, ws, nbytes = self._gn_table
        d0 = live[due[0]].discriminator
        cap, c, h, w = d0.memory.samples.shape
        iters = [int(v) for v in d0.update_iters]
        arr = (ctypes.c_int * len(iters))(*iters)
        lib().gn_update_batched(ptr(table), len(due), 1, cap, c, h, w, arr, len(iters), float(d0.filter_reg[-1]),
                                float(d0.precond[-1]), float(d0.direction_forget_factor), int(d0.min_px), 0, ptr(ws), nbytes, stream())

    def _ensure_gn_table(self, live, due):
        """(Re)build the device pointer table and the workspace of the batched filter update when the object set changes."""
        import ctypes
        from .._lib import lib, ptr, stream
        key = tuple((live[k].object_id, live[k].discriminator.filter.weight.data_ptr(), live[k].discriminator.memory.samples.data_ptr(),
                     live[k].discriminator.update_optimizer.cg_state.data_ptr()) for k in due) + (self._counts.data_ptr(),)
        if getattr(self, "_gn_table", None) is None or self._gn_table[0] != key:
            rows = [[], [], [], [], [], [], [], []]
            for k in due:
                d = live[k].discriminator
                m = d.memory
                for r, v in zip(rows, (m.samples, m.stencil, m.uty, m.weights, d.filter.weight, d.update_optimizer.cg_state)):
                    r.append(v.data_ptr())
                rows[6].append(self._counts.data_ptr() + 4 * k)
                rows[7].append(m.split.data_ptr())
            flat = [v for r in rows for v in r]
            # the table and the workspace keep their addresses when the object count is unchanged (they are part of a
            # captured block graph); only the contents are rewritten
            old = getattr(self, "_gn_table_buf", None)
            d0 = live[due[0]].discriminator
            cap, c, h, w = d0.memory.samples.shape
            nbytes = len(due) * lib().gn_update_workspace(cap, c, h, w)
            if old is not None and old[0].shape == (8, len(due)) and old[2] >= nbytes and old[0].device == self._counts.device:
                table, ws = old[0], old[1]
            else:
                table = torch.empty((8, len(due)), dtype=torch.int64, device=self._counts.device)
                ws = torch.empty(nbytes // 4, device=table.device, dtype=torch.float32)
                self._gn_table_buf = (table, ws, nbytes)
            lib().fill_i64(ptr(table), (ctypes.c_int64 * len(flat))(*flat), len(flat), stream())   # no synchronising H2D
            self._gn_table = (key, table, ws, nbytes)
