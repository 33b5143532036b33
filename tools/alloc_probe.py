import os, sys, time, gc
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
from frtm_vos_b200.model.memory import Memory
dev = "cuda:0"
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.time(); m = Memory(80, (96, 30, 54), (1, 480, 854), dev, 0.1); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    print("Memory() host %.2f ms, +sync %.2f ms, device allocs %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, torch.cuda.memory_stats()["num_device_alloc"]))
    del m
trk = build_tracker("resnet18", (480, 854), dev)
seq = synth.SyntheticSequence(num_objects=3, num_frames=12, size=(480, 854), seq_id=1)
seq.preload(dev)
for rep in range(3):
    a0 = torch.cuda.memory_stats()["num_device_alloc"]; f0 = torch.cuda.memory_stats()["num_device_free"]
    t0 = time.time(); trk.run_sequence(seq); t1 = time.time()
    print("run %d: %.1f ms, device allocs +%d frees +%d, reserved %.1f GB, gc objects %d" % (rep, (t1 - t0) * 1e3,
          torch.cuda.memory_stats()["num_device_alloc"] - a0, torch.cuda.memory_stats()["num_device_free"] - f0,
          torch.cuda.memory_reserved() / 1e9, gc.collect()))
