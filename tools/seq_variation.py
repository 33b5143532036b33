"""Step time of the bench workload for the sequences the ranks of an N-GPU run take (seq_id = 1 + rank), measured one after the
other on ONE GPU: separates "rank r has a slower sequence" from cross-process contention in the weak-scaling numbers."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
dev = "cuda:0"
size = (480, 854)
trk = build_tracker("resnet18", size, dev)
orig_init = trk.initialize
acc = dict(init=0.0, init_host=0.0)
def timed_init(*a):
    torch.cuda.synchronize(); t0 = time.time(); r = orig_init(*a); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    acc["init"] += t2 - t0; acc["init_host"] += t1 - t0; return r
for sid in range(1, 9):
    seq = synth.SyntheticSequence(num_objects=3, num_frames=65, size=size, seq_id=sid)
    areas = [float((seq[0][1] == o).float().mean()) for o in seq.obj_ids]
    seq.preload(dev)
    trk.initialize = orig_init
    for _ in range(2):
        trk.run_sequence(seq)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        trk.run_sequence(seq)
    torch.cuda.synchronize()
    plain = (time.time() - t0) / 3
    trk.initialize = timed_init
    acc.update(init=0.0, init_host=0.0)
    trk.run_sequence(seq)
    print("seq %d: step %.1f ms | init %.1f ms (host %.1f) | object areas %s" % (
        sid, plain * 1e3, acc["init"] * 1e3, acc["init_host"] * 1e3, " ".join("%.3f" % a for a in areas)), flush=True)
