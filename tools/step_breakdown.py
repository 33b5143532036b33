"""Wall-clock breakdown of one sequence of a BASELINE config (synchronising after each phase; profiling aid only):
python tools/step_breakdown.py [2|3|5]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
dev = "cuda:0"
sys.path.insert(0, ROOT)
import bench as B
cfg = B.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
size = cfg["size"]
trk = build_tracker(cfg["arch"], size, dev, fast=cfg["fast"], memory_size=cfg["memory"])
seq = synth.SyntheticSequence(num_objects=cfg["objects"], num_frames=cfg["frames"], size=size, seq_id=1)
seq.preload(dev)
for _ in range(2):
    trk.run_sequence(seq)
torch.cuda.synchronize()
orig_init, orig_block = trk.initialize, trk._track_block
acc = dict(init=0.0, init_host=0.0, block=0.0, block_host=0.0, nblocks=0)
def timed_init(*a):
    torch.cuda.synchronize(); t0 = time.time(); r = orig_init(*a); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    acc["init"] += t2 - t0; acc["init_host"] += t1 - t0; return r
def timed_block(imgs):
    torch.cuda.synchronize(); t0 = time.time(); r = orig_block(imgs); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    acc["block"] += t2 - t0; acc["block_host"] += t1 - t0; acc["nblocks"] += 1; return r
trk.initialize, trk._track_block = timed_init, timed_block
t0 = time.time(); trk.run_sequence(seq); total = time.time() - t0
print(cfg["name"])
print("total %.1f ms | init %.1f ms (host part %.1f) | %d blocks %.1f ms (host part %.1f) | other %.1f ms" % (
    total * 1e3, acc["init"] * 1e3, acc["init_host"] * 1e3, acc["nblocks"], acc["block"] * 1e3, acc["block_host"] * 1e3,
    (total - acc["init"] - acc["block"]) * 1e3))
