"""Host-side timeline of back-to-back sequences with the next sequence's augmentation prepared behind the tracking:
python tools/prefetch_timeline.py [2|3|5]   (no synchronisation inside the sequence; wall clock of the host calls)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
import bench as B
dev = "cuda:0"
cfg = B.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
size = cfg["size"]
trk = build_tracker(cfg["arch"], size, dev, fast=cfg["fast"], memory_size=cfg["memory"])
seq = synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=cfg["frames"], size=size, seq_id=1)
seq.preload(dev)
for mode in (False, True):
    trk.prefetch_next = mode
    trk._prefetched.clear()
    for _ in range(3):
        trk.run_sequence(seq, next_sequence=seq)
    torch.cuda.synchronize()
    marks = []
    oi, ob, op = trk.initialize, trk._track_block, trk.prefetch_init
    def ti(*a, **k):
        t0 = time.perf_counter(); r = oi(*a, **k); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        marks.append(("initialize host %.2f ms, +%.2f ms until the device is idle" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))); return r
    def tb(imgs):
        t0 = time.perf_counter(); r = ob(imgs); marks.append("block host %.2f ms" % ((time.perf_counter() - t0) * 1e3)); return r
    def tp(*a, **k):
        t0 = time.perf_counter(); r = op(*a, **k); marks.append("prefetch_init submit %.2f ms" % ((time.perf_counter() - t0) * 1e3)); return r
    trk.initialize, trk._track_block, trk.prefetch_init = ti, tb, tp
    t0 = time.perf_counter()
    trk.run_sequence(seq, next_sequence=seq)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    trk.initialize, trk._track_block, trk.prefetch_init = oi, ob, op
    print("== %s prefetch=%s: sequence %.1f ms" % (cfg["name"], mode, total))
    for m in marks:
        print("   ", m)
