"""Profiling target: one warm sequence, then one sequence between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off ...`).  Usage: python tools/profile_step.py [--config 2] [--frames 17]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--objects", type=int, default=3)
    ap.add_argument("--frames", type=int, default=17)
    ap.add_argument("--size", default="480x854")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--what", default="sequence", choices=["sequence", "track", "update", "block", "init"])
    a = ap.parse_args()
    from quick_run import build_tracker
    from frtm_vos_b200 import synth
    size = tuple(int(v) for v in a.size.split("x"))
    dev = "cuda:0"
    trk = build_tracker(a.arch, size, dev, fast=not a.full)
    seq = synth.SyntheticSequence(num_objects=a.objects, num_frames=a.frames, size=size, seq_id=1)
    seq.preload(dev)
    trk.run_sequence(seq)
    torch.cuda.synchronize()
    img = seq[a.frames - 1][0]
    d = trk.targets[1].discriminator
    torch.cuda.profiler.start()
    if a.what == "sequence":
        trk.run_sequence(seq)
    elif a.what == "track":
        trk.track(img)
    elif a.what == "block":
        imgs = [seq[a.frames - 1 - j][0] for j in range(8)]
        rem = (-d.frame_num) % d.train_skipping
        torch.cuda.profiler.stop()
        if rem:
            trk._track_block(imgs[:rem])
        trk._track_block(imgs)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trk._track_block(imgs)
    elif a.what == "init":
        from frtm_vos_b200.model.discriminator import Discriminator
        m5 = None
        im5, m5 = trk.augment(img, (seq.ground_truth(a.frames - 1) == 1).byte().to(dev))
        _, nh, _ = trk.feature_extractor.forward_split(im5, (), ("layer4",), upto="layer4")
        dd = Discriminator(**trk.disc_params)
        torch.cuda.synchronize()
        dd.init(None, m5, x_nhwc=nh["layer4"])
    else:
        d.update_optimizer.run(d.update_iters)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
