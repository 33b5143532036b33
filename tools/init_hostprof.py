"""cProfile of Tracker.initialize on a warm tracker with the next-sequence prefetch active (host side of what is left of an
initialisation): python tools/init_hostprof.py [2|3]"""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
import bench as B
dev = "cuda:0"
cfg = B.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
trk = build_tracker(cfg["arch"], cfg["size"], dev, fast=cfg["fast"], memory_size=cfg["memory"])
seq = synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=cfg["frames"], size=cfg["size"], seq_id=1)
seq.preload(dev)
for _ in range(3):
    trk.run_sequence(seq, next_sequence=seq)
torch.cuda.synchronize()
pr = cProfile.Profile()
oi = trk.initialize
def prof_init(*a, **k):
    pr.enable(); r = oi(*a, **k); pr.disable(); return r
trk.initialize = prof_init
for _ in range(3):
    trk.run_sequence(seq, next_sequence=seq)
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
