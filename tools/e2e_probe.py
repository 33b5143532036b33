"""Where the end-to-end step differs from the device-resident one: the four combinations of frames on the host / on the
device and label maps read back / left on the device (config 2).  python tools/e2e_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
dev = "cuda:0"
cfg = B.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
trk, dp = B.build_tracker(cfg, dev)
H, W = cfg["size"]
seq = B.sequence_for(cfg, 1)
hseq = B.HostSequence(seq)
seq.preload(dev)
pinned = torch.empty((cfg["frames"], H, W), dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for host_in in (False, True, False, True):
    for host_out in (False, True):
        s = hseq if host_in else seq
        for _ in range(3):
            trk.run_sequence(s, next_sequence=s, host_labels=pinned if host_out else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            trk.run_sequence(s, next_sequence=s, host_labels=pinned if host_out else None)
            flush.zero_()
        e1.record()
        torch.cuda.synchronize()
        print("frames on %s, labels %s: %.2f ms per sequence" % ("host  " if host_in else "device", "to pinned host" if host_out else "left on device", e0.elapsed_time(e1) / 8))
