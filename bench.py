#!/usr/bin/env python
"""bench.py — frames/sec of the FRTM per-frame inference hot path on synthetic DAVIS-shaped video.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|5]

One "step" = one full `Tracker.run_sequence` over a synthetic sequence (per-object initialisation included, exactly
as the reference defines fps, model/tracker.py:130,159-161).  Default workload = BASELINE.json configs[1]:
ResNet18 --fast, 3 objects, 65 frames of 480x854, one sequence per GPU (weak scaling, sequences are independent;
the only collective is the end-of-step NCCL all_gather of the uint8 label maps).

Prints ONE JSON line (rank 0).  `value` = frames/s with frames resident in HBM; `e2e` = the same through the public
API with pinned HOST frames (H2D of every frame and D2H of every label map inside the timed region);
`roofline` = the GN/CG operator kernel sequence (one A·p over the frame memory) against the measured HBM peak;
`roofline_conv` = conv-path algorithmic FLOP/s against the measured bf16 tensor peak; `cpu_baseline` = the CPU oracle
(port of the reference path) timed on this box's host cores on a bounded sample.  `--impl reference` runs only that
CPU arm and prints it as its own line.
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
    2: dict(arch="resnet18", fast=True, objects=3, frames=65, size=(480, 854), memory=80,
            name="rn18-fast/3obj/65f/480x854"),
    3: dict(arch="resnet101", fast=False, objects=5, frames=69, size=(480, 854), memory=80,
            name="rn101-full/5obj/69f/480x854"),
    5: dict(arch="resnet101", fast=False, objects=10, frames=65, size=(720, 1280), memory=32,
            name="rn101-full/10obj/65f/720x1280/mem32"),
}
# algorithmic conv GFLOP per frame (measured by hooking the reference)
CONV_GFLOP = {("resnet18", 480): (29.82, 23.07 + 0.082), ("resnet101", 480): (128.63, 24.25 + 0.321),
              ("resnet18", 720): (66.96, 51.73 + 0.183), ("resnet101", 720): (287.08, 54.38 + 0.714)}


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
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
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(cfg, sample_frames, steps, warmup):
    """The reference's CPU path (oracle port, all host threads) on a bounded sample of the same workload."""
    from oracle import frtm_ref as R
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    import golden_inputs as GI
    torch.set_num_threads(os.cpu_count())
    bb = synth.backbone_state_dict(cfg["arch"], size=cfg["size"])
    seg = synth.segnet_state_dict(cfg["arch"])
    dp = GI.oracle_disc_params(disc_params(cfg, "cpu"))
    seq = synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=sample_frames, size=cfg["size"], seq_id=1)
    trk = R.TrackerRef(bb, cfg["arch"], seg, dp, ImageAugmenter(GI.AUG_PARAMS).augment_first_frame, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        trk.run_sequence(seq)
        if i >= warmup:
            times.append(time.time() - t0)
    dt = sum(times) / len(times)
    return dict(value=sample_frames / dt, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                sample="first %d frames (incl. %d object inits) of %s, oracle/frtm_ref.py on CPU, %d step(s)" % (
                    sample_frames, cfg["objects"], cfg["name"], len(times))), dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample-frames", type=int, default=9)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if a.impl == "reference":
        if rank != 0:
            return
        k = max(1, min(a.steps, 2))
        cb, dt = cpu_reference_arm(cfg, a.cpu_sample_frames, k, min(a.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": "frames/sec (480p, multi-object)", "value": cb["value"], "unit": "frames/s",
            "n_gpus": a.gpus, "steps": k, "warmup": min(a.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample_frames": a.cpu_sample_frames},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    # host side per rank = OpenCV inpaint + kernel launches: do not oversubscribe the host cores with N ranks
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
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

    from frtm_vos_b200 import synth, ops
    from frtm_vos_b200._lib import lib
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

    # every rank tracks its own, different sequence (object/sequence-sharded, no collective inside the frame loop)
    seq = synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=cfg["frames"], size=cfg["size"], seq_id=1 + rank)
    host_frames = [seq[t] for t in range(len(seq))]

    class HostSequence:  # same protocol, frames in pinned host memory -> H2D happens inside run_sequence
        name, obj_ids, frame_names = seq.name, seq.obj_ids, seq.frame_names

        def __init__(self):
            self.items = [(im.pin_memory(), (lb.pin_memory() if torch.is_tensor(lb) else lb), ids) for im, lb, ids in host_frames]

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return self.items[i]

    seq.preload(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(sequence, read_back=False):
        outs, _ = trk.run_sequence(sequence)
        labels = torch.stack([o.reshape(cfg["size"]) for o in outs])
        if world > 1:
            from frtm_vos_b200.parallel import gather_label_maps
            gather_label_maps(labels.unsqueeze(0), world)        # end-of-batch gather of the label maps (NCCL/NVLink)
        if read_back:
            return labels.cpu()
        return labels

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step(seq)
        flush.zero_()
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L = lib()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step(seq)
        flush.zero_()          # L2 flush between timed iterations (256 MiB > 126 MB L2)
    e1.record()
    sync()
    launches = L.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # e2e: pinned host frames in, label maps out, through the public run_sequence API
    hseq = HostSequence()
    step(hseq, True)
    sync()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_e2e = max(1, min(a.steps, 3))
    e2.record()
    for _ in range(k_e2e):
        step(hseq, True)
        flush.zero_()
    e3.record()
    sync()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- roofline of the CG operator on this rank's real frame memories (all objects, one batched update) -------
    pk = peaks()
    live = [trk.targets[o] for o in seq.obj_ids]
    discs = [t.discriminator for t in live]
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
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    r0.record()
    for _ in range(reps):
        trk._batched_gn_update(live, due)
    r1.record()
    torch.cuda.synchronize()
    for dd, (f0, s0) in zip(discs, saved):
        dd.filter.weight.data.copy_(f0); dd.update_optimizer.cg_state.copy_(s0)
    ms_update = r0.elapsed_time(r1) / reps
    M = sum(Ms)
    ap_bytes = M * 4 * (c * h * w + 9 * h * w)                  # form S, one A.p over all objects
    rhs_bytes = M * 4 * (c * h * w + 10 * h * w)
    launches_per_update = n_cg + 1          # the operator kernel's last CTA per object runs the CG vector step itself
    achieved = (rhs_bytes + n_cg * ap_bytes) / (ms_update * 1e-3) / 1e9
    # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/r01_ncu_full_summary.md): 211 MB read for
    # 3 objects x 69 samples = 1.02 MB per active sample — 1.5x the algorithmic bytes: the second (reversed) pass over a
    # sample finds about half of it still in L2 when 207 samples stream at once; scaled to this launch.
    traffic = 211.0e6 / 207 * M if (c, h, w) == (96, 30, 54) else None
    roofline = dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"], traffic=traffic,
                    kernel="gn_apply_tc_kernel (tcgen05 operator over the operator images, CG vector step fused into its tail) "
                           "inside one batched filter update (RHS + %d x A.p, stencil form S; %d objects, M=%s active samples "
                           "of %d, sample = %dx%dx%d as split fp16 planes = fp32 bytes)" % (n_cg, len(live), Ms, cap, c, h, w),
                    ms=ms_update, launches=launches_per_update, algorithmic_bytes_per_launch=ap_bytes, peak_source=pk["src"],
                    note="working set %.0f MB vs 126 MB L2" % (M * 4 * (c + 10) * h * w / 1e6))

    # ---- conv path: algorithmic FLOP/s of one 8-frame block (the unit run_sequence executes) -------------------------
    nblk = trk.max_block
    imgs = [seq[len(seq) - 1 - j][0] for j in range(nblk)]
    rem = (-trk.targets[seq.obj_ids[0]].discriminator.frame_num) % trk.targets[seq.obj_ids[0]].discriminator.train_skipping
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
                                "update; tcgen05 split-fp16 convs execute hi*hi + hi*lo + lo*hi per algorithmic MAC (executed tensor FLOPs = 3x)" % cfg["objects"],
                         ms_per_frame=ms_track, algorithmic_gflop_per_frame=gb + cfg["objects"] * go, peak_source=pk["src"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = cfg["frames"]
    value = world * a.steps * frames / (ms * 1e-3)
    e2e_v = world * k_e2e * frames / (ms_e2e * 1e-3)
    h2d = frames * 3 * cfg["size"][0] * cfg["size"][1] + cfg["size"][0] * cfg["size"][1]
    d2h = frames * cfg["size"][0] * cfg["size"][1]
    out = {
        "metric": "frames/sec (480p, multi-object)", "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sequences_per_gpu": 1, "objects": cfg["objects"], "frames": frames,
                   "init_iters": list(dp["init_iters"]), "update_iters": list(dp["update_iters"]), "memory_size": cfg["memory"],
                   "l2": "256 MiB buffer written between steps", "parallelism": "sequence-sharded x%d, end-of-step all_gather of labels" % world,
                   "host": "%d cores, wait policy %s" % (cores, wait_policy)},
        "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k_e2e},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_conv": roofline_conv,
    }
    if not a.no_cpu_baseline and world == 1:
        cb, _ = cpu_reference_arm(cfg, a.cpu_sample_frames, 1, 0)
        out["cpu_baseline"] = cb
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
