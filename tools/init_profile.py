"""cProfile of the main thread during the first-frame initialisation of a warm sequence (GPU box): python tools/init_profile.py"""
import cProfile, os, pstats, sys, time, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
dev = "cuda:0"
size = (480, 854)
trk = build_tracker("resnet18", size, dev)
seq = synth.SyntheticSequence(num_objects=3, num_frames=9, size=size, seq_id=1)
seq.preload(dev)
for _ in range(3):
    trk.run_sequence(seq)
torch.cuda.synchronize()
orig = trk.initialize
pr = cProfile.Profile()
times = []
def prof_init(*a):
    torch.cuda.synchronize(); t0 = time.time()
    pr.enable(); r = orig(*a); pr.disable()
    t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    times.append((t1 - t0, t2 - t0)); return r
trk.initialize = prof_init
for _ in range(3):
    trk.run_sequence(seq)
print("init host/total ms:", [(round(a * 1e3, 1), round(b * 1e3, 1)) for a, b in times])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
