"""Host-side timeline of Tracker.initialize on a warm tracker (GPU box): python tools/init_timeline.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_run import build_tracker
from frtm_vos_b200 import synth
from frtm_vos_b200.model import tracker as T, discriminator as D
dev = "cuda:0"
size = (480, 854)
trk = build_tracker("resnet18", size, dev)
seq = synth.SyntheticSequence(num_objects=3, num_frames=9, size=size, seq_id=1)
seq.preload(dev)
for _ in range(3):
    trk.run_sequence(seq)
torch.cuda.synchronize()
marks = []
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); marks.append((label, (time.perf_counter() - t0) * 1e3)); return r
    setattr(obj, name, g)
wrap(T, "TargetObject", "TargetObject()")
wrap(trk.feature_extractor, "forward_split", "forward_split")
orig_init = D.Discriminator.init
def dinit(self, *a, **k):
    t0 = time.perf_counter(); r = orig_init(self, *a, **k); marks.append(("disc.init", (time.perf_counter() - t0) * 1e3)); return r
D.Discriminator.init = dinit
from frtm_vos_b200.model import augmenter as A
for nm in ("cut_and_inpaint", "mask_center_bbox", "draw_specs", "target_locations"):
    wrap(A, nm, "  aug." + nm)
wrap(trk.augment.__self__ if hasattr(trk.augment, "__self__") else trk, "_warp_masks_device", "  aug._warp_masks_device") if hasattr(getattr(trk.augment, "__self__", None), "_warp_masks_device") else None
if hasattr(getattr(trk.augment, "__self__", None), "_render_device"):
    wrap(trk.augment.__self__, "_render_device", "  aug._render_device")
orig_aug = trk.augment
def aug(*a, **k):
    t0 = time.perf_counter(); r = orig_aug(*a, **k); marks.append(("augment(thread)", (time.perf_counter() - t0) * 1e3)); return r
trk.augment = aug
trk._augment_takes_rng = True
oi = trk.initialize
def init(*a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = oi(*a); t1 = time.perf_counter(); torch.cuda.synchronize()
    marks.append(("initialize host", (t1 - t0) * 1e3)); marks.append(("initialize total", (time.perf_counter() - t0) * 1e3)); return r
trk.initialize = init
for _ in range(2):
    marks.clear()
    trk.run_sequence(seq)
for m in marks:
    if m[0] != "forward_split" or True:
        print("%-20s %.2f ms" % m)
