"""Per-frame wall-clock (synchronised) of a sequence + cProfile of the host side of initialize()."""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
dev = "cuda:0"
size = (480, 854)
trk = build_tracker("resnet18", size, dev)
seq = synth.SyntheticSequence(num_objects=3, num_frames=20, size=size, seq_id=1)
seq.preload(dev)
trk.run_sequence(seq)
torch.cuda.synchronize()
# manual loop with per-frame timing
trk.targets = dict(); trk.current_frame = 0; trk._stack = None; trk._fbuf = None; trk._gn_table = None
trk.object_ids = seq.obj_ids
trk._lut = torch.tensor([0] + list(seq.obj_ids), dtype=torch.uint8, device=dev)
pr = cProfile.Profile()
for i in range(len(seq)):
    image, labels, new = seq[i]
    torch.cuda.synchronize(); t0 = time.time()
    had = len(trk.targets) > 0
    if new:
        pr.enable(); trk.initialize(image, labels.to(dev), new); pr.disable()
    if had:
        trk.track(image)
    t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    print("frame %2d host %.2f ms  +gpu drain %.2f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    trk.current_frame += 1
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:4500])
