#!/usr/bin/env python
"""bench.py — frames/sec of the FRTM per-frame inference hot path on synthetic DAVIS-shaped video.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--also 3,5]
                    [--impl b200|reference|reference-cuda]

One "step" = one full `Tracker.run_sequence` over every synthetic sequence of this rank (per-object initialisation
included, exactly as the reference defines fps, model/tracker.py:130,159-161).  Default workload = BASELINE.json
configs[1]: ResNet18 --fast, 3 objects, 65 frames of 480x854, one sequence per GPU (weak scaling: sequences are
independent; the only collective is the end-of-step NCCL all_gather of the uint8 label maps).  `--config 4` is
BASELINE configs[3]: 64 sequences x 4 objects x 33 frames (rn101), sharded over the ranks by `assign_sequences`
(strong scaling).

Rank 0 prints ONE JSON line.  `value` = frames/s with frames resident in HBM; `e2e` = the same through the public API with
pinned HOST frames (H2D of every frame and D2H of every label map inside the timed region); `roofline` = the GN/CG
operator (one filter update = RHS + n_cg operator applications over the frame memory) against the measured HBM peak;
`roofline_conv` = conv-path algorithmic FLOP/s against the measured bf16 tensor peak.  At N = 1 the line also carries
  `cpu_baseline`   the CPU oracle (port of the reference path) timed on this box's host cores over the FULL sequence,
  `parity`         labels / logits of the CUDA path against that oracle run on the first 9 frames at fixed target-model state,
  `reference_cuda` the same oracle (plain PyTorch: cuDNN convs, autograd double-backward GN/CG, cudnn.benchmark=True as
                   evaluate.py:15-16) on cuda:0 with TF32 on (the reference's default) and off — the CUDA/PyTorch build the
                   north-star names as the bar,
  `other_configs`  the BASELINE configs 3 and 5 (value, e2e, both rooflines) measured in the same process.
`--impl reference` runs only the CPU arm (full sequence per step, as many of the K steps as fit FRTM_REF_BUDGET_S,
default 200 s); `--impl reference-cuda` only the PyTorch-on-GPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if int(os.environ.get("WORLD_SIZE", 1)) > 1:
    # several ranks share the host: idle OpenMP workers of the few CPU-side torch ops must sleep, not spin
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import torch  # noqa: E402

CONFIGS = {
    2: dict(arch="resnet18", fast=True, objects=3, frames=65, size=(480, 854), memory=80, sequences=0,
            name="rn18-fast/3obj/65f/480x854"),
    3: dict(arch="resnet101", fast=False, objects=5, frames=69, size=(480, 854), memory=80, sequences=0,
            name="rn101-full/5obj/69f/480x854"),
    4: dict(arch="resnet101", fast=False, objects=4, frames=33, size=(480, 854), memory=80, sequences=64,
            name="rn101-full/64seq x 4obj/33f/480x854"),
    5: dict(arch="resnet101", fast=False, objects=10, frames=65, size=(720, 1280), memory=32, sequences=0,
            name="rn101-full/10obj/65f/720x1280/mem32"),
}
# algorithmic conv GFLOP per frame (measured by hooking the reference)
CONV_GFLOP = {("resnet18", 480): (29.82, 23.07 + 0.082), ("resnet101", 480): (128.63, 24.25 + 0.321),
              ("resnet18", 720): (66.96, 51.73 + 0.183), ("resnet101", 720): (287.08, 54.38 + 0.714)}
METRIC = "frames/sec (480p, multi-object)"
PARITY_FRAMES = 9          # frame 0 (init) + 8 tracked frames, the last of which triggers the first filter update


def disc_params(cfg, dev):
    import golden_inputs as GI
    from frtm_vos_b200 import synth
    C = synth.backbone_out_channels(cfg["arch"])["layer4"]
    return GI.disc_params(C, init_iters=(5, 10, 10, 10) if cfg["fast"] else (5, 10, 10, 10, 10),
                          update_iters=(5,) if cfg["fast"] else (10,), memory_size=cfg["memory"], device=dev)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


def workload_config(cfg, dp, world, cores, wait_policy):
    """The `config` object of the JSON line — identical for the b200 and reference arms of the same workload."""
    return {"workload": cfg["name"], "objects": cfg["objects"], "frames": cfg["frames"],
            "sequences": cfg["sequences"] or "1 per GPU", "init_iters": list(dp["init_iters"]),
            "update_iters": list(dp["update_iters"]), "memory_size": cfg["memory"],
            "l2": "256 MiB buffer written between device steps",
            "parallelism": "sequence-sharded, end-of-step all_gather of labels"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 20 ms from a thread (the counters
    nvidia-smi prints), or — when pynvml cannot open the device — an `nvidia-smi -lms` child process."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.max_mhz = None, [], None
        self._halt = threading.Event()

    def _start_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = None
        try:                                               # CUDA_VISIBLE_DEVICES may renumber: look the device up by UUID
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            for u in (uuid, uuid.encode()):
                try:
                    h = N.nvmlDeviceGetHandleByUUID(u)
                    break
                except Exception:
                    h = None
        except Exception:
            h = None
        if h is None:
            h = N.nvmlDeviceGetHandleByIndex(self.index)

        def reasons():
            try:
                return int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                return int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))

        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)       # probes: fail here, not in the thread
        reasons()

        def loop():
            while not self._halt.is_set():
                try:
                    self.samples.append((float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), reasons()))
                except Exception:
                    pass
                self._halt.wait(0.02)

        self.nvml = threading.Thread(target=loop, daemon=True)
        self.nvml.start()

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self._halt.set()
            self.nvml.join(timeout=1.0)
            sm = sorted(v for v, _ in self.samples)
            bits = 0
            for _, r in self.samples:
                bits |= r
            return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=self.max_mhz,
                        reasons=[n for b, n in self.BITS if bits & b], samples=len(sm), source="nvml")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                    source="nvidia-smi")


# ------------------------------------------------------------------------------------------------------------------
# reference arms: the oracle port of the reference path (plain PyTorch) on the host cores or on cuda:0
# ------------------------------------------------------------------------------------------------------------------
def oracle_tracker(cfg, device, hooks=None):
    from oracle import frtm_ref as R
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    import golden_inputs as GI
    bb = synth.backbone_state_dict(cfg["arch"], size=cfg["size"])
    seg = synth.segnet_state_dict(cfg["arch"])
    dp = GI.oracle_disc_params(disc_params(cfg, "cpu"))
    # first-frame augmentation rendered on the host with OpenCV (device_render=False): none of this repo's kernels run
    aug = ImageAugmenter(GI.AUG_PARAMS, device_render=False).augment_first_frame
    return R.TrackerRef(bb, cfg["arch"], seg, dp, aug, device, hooks=hooks)


def sequence_for(cfg, seq_id, frames=None):
    from frtm_vos_b200 import synth
    return synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=frames or cfg["frames"], size=cfg["size"], seq_id=seq_id)


def cpu_reference_arm(cfg, steps, warmup, frames=0, budget_s=None, hooks=None):
    """The reference's CPU path (oracle port, all host threads) over the same sequence the b200 arm tracks (frames = 0:
    all of it).  Runs `warmup` + up to `steps` passes, stopping early when `budget_s` is spent.  -> (cpu_baseline, s/step, n, labels)"""
    torch.set_num_threads(os.cpu_count())
    n_frames = frames or cfg["frames"]
    seq = sequence_for(cfg, 1, n_frames)
    trk = oracle_tracker(cfg, "cpu", hooks)
    times, out, t_start = [], None, time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        out, _ = trk.run_sequence(seq)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.time() - t_start + dt > budget_s:
            break
    dt = sum(times) / len(times)
    what = "all %d frames" % n_frames if n_frames == cfg["frames"] else "first %d of %d frames" % (n_frames, cfg["frames"])
    return dict(value=n_frames / dt, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                sample="%s (incl. %d object inits) of %s, oracle/frtm_ref.py on CPU, %d timed pass(es)" % (
                    what, cfg["objects"], cfg["name"], len(times))), dt, len(times), out


def reference_cuda_arm(cfg, steps, warmup, dev="cuda:0"):
    """The oracle port on cuda:0 = what the reference's own CUDA/PyTorch build executes: cuDNN convolutions, ATen
    bilinear / softmax, autograd double-backward GN/CG, `cudnn.benchmark = True` (evaluate.py:15-16); with TF32 convs (the
    torch default the reference runs with) and without (the precision this repo's path is held to)."""
    seq = sequence_for(cfg, 1)
    seq.preload(dev)
    res = {}
    torch.backends.cudnn.benchmark = True
    for name, tf32 in (("tf32", True), ("fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False            # torch default; the reference does not change it
        trk = oracle_tracker(cfg, dev)
        for _ in range(warmup):
            trk.run_sequence(seq)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            trk.run_sequence(seq)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[name] = dict(value=cfg["frames"] / (ms * 1e-3), unit="frames/s", ms_per_step=ms, steps=steps, warmup=warmup)
        del trk
        torch.cuda.empty_cache()
    res["what"] = ("oracle/frtm_ref.py (plain PyTorch port of the reference path: cuDNN convs, autograd GN/CG) on cuda:0, "
                   "cudnn.benchmark=True, frames resident on the device, %s" % cfg["name"])
    return res


# ------------------------------------------------------------------------------------------------------------------
# the b200 arm
# ------------------------------------------------------------------------------------------------------------------
class HostSequence:
    """Same protocol as the synthetic sequence, frames in pinned host memory -> H2D happens inside run_sequence."""

    def __init__(self, seq):
        self.name, self.obj_ids, self.frame_names = seq.name, seq.obj_ids, seq.frame_names
        self.items = [(im.pin_memory(), (lb.pin_memory() if torch.is_tensor(lb) else lb), ids)
                      for im, lb, ids in (seq[t] for t in range(len(seq)))]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def build_tracker(cfg, dev):
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200.model.seg_network import SegNetwork
    from frtm_vos_b200.model.tracker import Tracker
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    import golden_inputs as GI
    bb = synth.backbone_state_dict(cfg["arch"], size=cfg["size"])
    seg = synth.segnet_state_dict(cfg["arch"])
    dp = disc_params(cfg, dev)
    fe = ResnetFeatureExtractor(cfg["arch"], state_dict=bb).to(dev)
    refiner = SegNetwork(1, 64, {L: c for L, c in fe.get_out_channels().items() if L != "layer1"}, True)
    trk = Tracker(ImageAugmenter(GI.AUG_PARAMS), fe, dp, refiner, dev)
    trk.load_state_dict(seg)
    trk.to(dev)
    return trk, dp


def run_b200(cfg, steps, warmup, dev, rank, world, dist, with_clocks):
    """Times the workload on this rank; returns the fields of the JSON line that describe it (max over ranks applied)."""
    from frtm_vos_b200._lib import lib
    from frtm_vos_b200.parallel import assign_sequences, gather_label_maps
    trk, dp = build_tracker(cfg, dev)
    H, W = cfg["size"]
    if cfg["sequences"]:
        # BASELINE config 4: a fixed set of sequences sharded over the ranks (strong scaling)
        plan = assign_sequences([cfg["frames"] * cfg["objects"]] * cfg["sequences"], world)
        mine = [sequence_for(cfg, 1 + i) for i in plan[rank]]
        total_frames = cfg["sequences"] * cfg["frames"]
        scaling = "strong"
    else:
        # every rank tracks its own, different sequence (no collective inside the frame loop)
        mine = [sequence_for(cfg, 1 + rank)]
        total_frames = world * cfg["frames"]
        scaling = "weak"
    hseqs = [HostSequence(s) for s in mine]
    for s in mine:
        s.preload(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    host_out = [torch.empty((cfg["frames"], H, W), dtype=torch.uint8).pin_memory() for _ in mine]

    # [sync] flag of run_sequence.  FRTM_BENCH_CHAIN=1 chains the sequences without a device synchronisation between them
    # (the next initialisation overlaps the last block): +0.7 % (frames in HBM) / +1.4 % (end to end) on B200; the default keeps
    # the reference's per-sequence synchronisation
    chain = [os.environ.get("FRTM_BENCH_CHAIN", "0") != "1"]

    def step(seqs, read_back=False):
        # the rank's sequences run back to back as in Tracker.run_dataset: while one is tracked, the host half of the next
        # one's first-frame initialisation (augmentation) is prepared in worker threads; the sequence after the last one
        # of a step is the first one of the next step (FRTM_PREFETCH_INIT=0 turns the overlap off).  With read_back the
        # label maps of every block go to pinned host memory on a copy stream while the next block is tracked.
        maps = []
        for k, s in enumerate(seqs):
            outs, _ = trk.run_sequence(s, next_sequence=seqs[(k + 1) % len(seqs)], host_labels=host_out[k] if read_back else None,
                                       sync=chain[0])
            maps.append(torch.stack([o.reshape(H, W) for o in outs]))
        labels = torch.stack(maps)                                 # (S,T,H,W) uint8
        if world > 1:
            gather_label_maps(labels, world)                       # end-of-batch gather of the label maps (NCCL/NVLink)
        return host_out if read_back else labels

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step(mine)
        flush.zero_()
    sync()
    sampler = ClockSampler(torch.cuda.current_device())
    if with_clocks:
        sampler.start()
    L = lib()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(mine)
        flush.zero_()          # L2 flush between timed iterations (256 MiB > 126 MB L2)
    e1.record()
    sync()
    launches = L.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if with_clocks else None

    # e2e: pinned host frames in, label maps out, through the public run_sequence API.  Its own warm-up: the host-frame
    # path allocates differently (per-frame upload buffers on the copy stream, the prefetch record of a host sequence),
    # and a first timed step that still grows the allocator's pools showed up as a 25-200 ms stall in one run out of four
    for _ in range(max(1, min(warmup, 3))):
        step(hseqs, True)
        flush.zero_()
    sync()
    k_e2e = max(1, min(steps, 8))
    k_np = max(1, min(steps, 3))
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(k_e2e):
        step(hseqs, True)
        flush.zero_()
    e3.record()
    sync()
    ms_e2e = e2.elapsed_time(e3)
    # the same device-resident steps with the cross-sequence overlap of the initialisation turned off
    trk.prefetch_next = False
    trk._drain_prefetch()
    trk._prefetched.clear()
    step(mine)
    sync()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(k_np):
        step(mine)
        flush.zero_()
    e5.record()
    sync()
    ms_np = e4.elapsed_time(e5)
    trk.prefetch_next = os.environ.get("FRTM_PREFETCH_INIT", "1") == "1"
    t = torch.tensor([ms, ms_e2e, ms_np], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_np = t.tolist()

    # ---- roofline of the GN/CG operator on this rank's real frame memories (all objects, one batched update) -------
    pk = peaks()
    seq = mine[-1]
    live = [trk.targets[o] for o in seq.obj_ids]
    discs = [tg.discriminator for tg in live]
    d = discs[0]
    cap, c, h, w = d.memory.samples.shape
    Ms = [int((dd.memory.weights > 0).sum().item()) for dd in discs]
    n_cg = sum(d.update_iters)
    saved = [(dd.filter.weight.data.clone(), dd.update_optimizer.cg_state.clone()) for dd in discs]
    trk._counts[:len(live)] = 1000            # open the device-side gate for the measurement
    due = list(range(len(live)))
    for _ in range(3):
        trk._batched_gn_update(live, due)
    torch.cuda.synchronize()
    reps, ms_up = 20, []
    for _ in range(reps):
        flush.zero_()                         # every update starts with a cold L2, as inside a sequence
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        trk._batched_gn_update(live, due)
        r1.record()
        torch.cuda.synchronize()
        ms_up.append(r0.elapsed_time(r1))
    for dd, (f0, s0) in zip(discs, saved):
        dd.filter.weight.data.copy_(f0); dd.update_optimizer.cg_state.copy_(s0)
    ms_update = sum(ms_up) / len(ms_up)
    M = sum(Ms)
    ap_bytes = M * 4 * (c * h * w + 9 * h * w)                  # form S, one A.p over all objects
    rhs_bytes = M * 4 * (c * h * w + 10 * h * w)
    n_launch = n_cg + 1                       # the operator kernel's last CTA per object runs the CG vector step itself
    achieved = (rhs_bytes + n_cg * ap_bytes) / (ms_update * 1e-3) / 1e9
    kind = L.gn_operator_kind(c, h, w)
    roofline = dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"],
                    traffic=measured_traffic(cfg, M),
                    kernel={4: "gn_apply_cl_kernel (every sample held on chip by a thread-block cluster through its three phases: one "
                               "HBM read per sample, mma.sync tiles, DSMEM halo exchange, persistent clusters, CG vector step fused)",
                            3: "gn_apply_mma_kernel (single pass over the operator images: sliding window, mma.sync tiles, CG vector "
                               "step fused into its tail)",
                            2: "gn_apply_tc_kernel (two-pass tcgen05 operator over the operator images, CG vector step fused)",
                            1: "gn_apply_kernel (CUDA cores)"}[int(kind)] +
                           " inside one batched filter update (RHS + %d x A.p, stencil form S; %d objects, M=%s active samples "
                           "of %d, sample = %dx%dx%d as split fp16 planes = fp32 bytes)" % (n_cg, len(live), Ms, cap, c, h, w),
                    ms=ms_update, launches=n_launch, algorithmic_bytes_per_launch=ap_bytes, peak_source=pk["src"],
                    note="every update timed after an L2 flush; working set %.0f MB vs 126 MB L2" % (M * 4 * (c + 10) * h * w / 1e6))

    # ---- conv path: algorithmic FLOP/s of one 8-frame block (the unit run_sequence executes) -------------------------
    nblk = trk.max_block
    imgs = [seq[len(seq) - 1 - j][0] for j in range(nblk)]
    d0 = trk.targets[seq.obj_ids[0]].discriminator
    rem = (-d0.frame_num) % d0.train_skipping
    if rem:
        trk._track_block(imgs[:rem])     # align to the next filter-update frame
    for _ in range(2):
        trk._track_block(imgs)           # frame_num stays aligned: every block ends on an update frame
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(4):
        trk._track_block(imgs)
    c1.record()
    torch.cuda.synchronize()
    ms_track = c0.elapsed_time(c1) / (4 * nblk)
    gb, go = CONV_GFLOP[(cfg["arch"], cfg["size"][0])]
    conv_tflops = (gb + cfg["objects"] * go) / ms_track
    roofline_conv = dict(bound="tensor", achieved=conv_tflops, peak=pk["tensor"], unit="TFLOP/s", frac=conv_tflops / pk["tensor"],
                         kernel="8-frame track block: backbone + %d x (project+filter+refinement) + merge + memory insert + filter "
                                "update; tcgen05 split-fp16 convs execute hi*hi + hi*lo + lo*hi per algorithmic MAC (executed tensor "
                                "FLOPs = 3x)" % cfg["objects"],
                         ms_per_frame=ms_track, algorithmic_gflop_per_frame=gb + cfg["objects"] * go, peak_source=pk["src"])

    trk._drain_prefetch()                     # the preparation started by the last step is never consumed: let it finish
    torch.cuda.synchronize()
    n_seq = len(mine)
    h2d = n_seq * (cfg["frames"] * 3 * H * W + H * W)
    d2h = n_seq * cfg["frames"] * H * W
    return dict(value=steps * total_frames / (ms * 1e-3), ms_per_step=ms / steps, scaling=scaling,
                e2e={"value": k_e2e * total_frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "steps": k_e2e},
                gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_conv=roofline_conv, dp=dp, trk=trk,
                pipeline={"what": "sequences run back to back (Tracker.run_dataset semantics): the host half of the next "
                                  "sequence's first-frame initialisation (OpenCV cut-out / inpaint, candidate masks) is prepared "
                                  "in worker threads behind the current sequence's tracking; every step does one such "
                                  "preparation and consumes one",
                          "without_overlap": {"value": k_np * total_frames / (ms_np * 1e-3), "unit": "frames/s",
                                              "ms_per_step": ms_np / k_np, "steps": k_np}})


def measured_traffic(cfg, M):
    """DRAM bytes (read + write) per operator launch from the committed `ncu --set full` capture of this configuration
    (profiles/r02_gn_traffic.json, written by tools/ncu_summary.py from the .ncu-rep), scaled by the active samples; null
    when no capture of the configuration is committed."""
    p = os.path.join(ROOT, "profiles", "r02_gn_traffic.json")
    if not os.path.isfile(p):
        return None
    ent = json.load(open(p)).get(cfg["name"])
    if not ent:
        return None
    return ent["dram_bytes_per_launch"] / ent["active_samples"] * M


def parity_check(cfg, dev, out_ref, dump):
    """CUDA path vs the oracle run that produced `cpu_baseline`, first PARITY_FRAMES frames at fixed target-model state."""
    import replay
    trk, _ = build_tracker(cfg, dev)
    seq = sequence_for(cfg, 1)
    seq.preload(dev)
    torch.manual_seed(11)
    out, got = replay.replay_on_device(trk, seq, dump, PARITY_FRAMES, dev)
    res = replay.compare(seq, cfg["size"], out, got, out_ref, dump, PARITY_FRAMES)
    res["how"] = ("oracle-replay: each object's post-init state (P, F, memory, CG state) taken from the CPU oracle run, then "
                  "frames 1..%d tracked by the CUDA path; tie_px = mismatches the oracle itself flips under +-1e-3 on the logits"
                  % (PARITY_FRAMES - 1))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--also", default="3,5", help="further BASELINE configs co-reported at N=1 ('' = none)")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="CPU arm: first n frames only (0 = the full sequence)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline / parity / reference_cuda")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)

    if a.impl == "reference":
        if rank != 0:
            return
        budget = float(os.environ.get("FRTM_REF_BUDGET_S", "200"))
        w = min(a.warmup, 1)
        cb, dt, n, _ = cpu_reference_arm(cfg, a.steps, w, a.cpu_sample_frames, budget_s=budget)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": a.gpus, "steps": n,
            "warmup": w, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, disc_params(cfg, "cpu"), 1, cores, "-"),
            "cpu_baseline": cb, "steps_requested": a.steps,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")

    if a.impl == "reference-cuda":
        if rank != 0:
            return
        rc = reference_cuda_arm(cfg, max(1, min(a.steps, 3)), max(1, min(a.warmup, 2)))
        print(json.dumps({
            "impl": "reference-cuda", "metric": METRIC, "value": rc["tf32"]["value"], "unit": "frames/s", "n_gpus": 1,
            "steps": rc["tf32"]["steps"], "warmup": rc["tf32"]["warmup"], "ms_per_step": rc["tf32"]["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 convs / f32", "data": "synthetic",
            "config": workload_config(cfg, disc_params(cfg, "cpu"), 1, cores, "-"), "reference_cuda": rc}))
        return

    # host side per rank = OpenCV inpaint + kernel launches: do not oversubscribe the host cores with N ranks
    torch.set_num_threads(max(1, cores // max(world, 1)))
    try:
        import cv2
        cv2.setNumThreads(max(1, cores // max(world, 1)))
    except Exception:
        pass
    from frtm_vos_b200.parallel import set_host_wait_policy
    wait_policy = os.environ.get("FRTM_HOST_WAIT", "yield")
    try:                                           # before torch creates the context: waiting host threads give their core away
        set_host_wait_policy(local, wait_policy)
    except (OSError, RuntimeError, KeyError) as e:
        print("bench.py: host wait policy left at the CUDA default (%s)" % (e,), file=sys.stderr)
        wait_policy = "default"
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    r = run_b200(cfg, a.steps, a.warmup, dev, rank, world, dist, with_clocks=(rank == 0))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    config = workload_config(cfg, r["dp"], world, cores, wait_policy)
    out = {
        "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config, "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"],
        "roofline": r["roofline"], "roofline_conv": r["roofline_conv"], "pipeline": r["pipeline"], "cpu_baseline": None,
        "host": "%d cores, wait policy %s" % (cores, wait_policy),
    }
    del r
    torch.cuda.empty_cache()
    if world == 1 and not a.no_cpu_baseline:
        import replay
        dump = {}
        sample = a.cpu_sample_frames or (0 if a.config == 2 else 17)      # the rn101 configs cost minutes per full CPU pass
        cb, _, _, out_ref = cpu_reference_arm(cfg, 1, 0, sample, hooks=replay.oracle_hooks(dump, PARITY_FRAMES - 1))
        out["cpu_baseline"] = cb
        try:
            out["parity"] = parity_check(cfg, dev, out_ref, dump)
        except Exception as e:                                            # noqa: BLE001 - the bench line must still print
            out["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.empty_cache()
        try:
            out["reference_cuda"] = reference_cuda_arm(cfg, 1, 1, dev)
        except Exception as e:                                            # noqa: BLE001
            out["reference_cuda"] = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.empty_cache()
        others = {}
        for k in [int(v) for v in a.also.split(",") if v.strip()]:
            if k == a.config or k not in CONFIGS:
                continue
            try:
                o = run_b200(CONFIGS[k], 3, 3, dev, 0, 1, None, with_clocks=True)
                others[CONFIGS[k]["name"]] = {
                    "value": o["value"], "unit": "frames/s", "ms_per_step": o["ms_per_step"], "steps": 3, "warmup": 3,
                    "e2e": o["e2e"], "gpu_launches": o["gpu_launches"], "clocks": o["clocks"], "roofline": o["roofline"],
                    "roofline_conv": o["roofline_conv"], "pipeline": o["pipeline"],
                    "config": workload_config(CONFIGS[k], o["dp"], 1, cores, wait_policy)}
                del o
            except Exception as e:                                        # noqa: BLE001
                others[CONFIGS[k]["name"]] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
        out["other_configs"] = others
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
