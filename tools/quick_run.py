"""Quick end-to-end run of the CUDA tracker on a synthetic sequence with per-stage CUDA-event timings."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_tracker(arch, size, dev, fast=True, memory_size=80):
    from frtm_vos_b200 import synth
    from frtm_vos_b200.model.feature_extractor import ResnetFeatureExtractor
    from frtm_vos_b200.model.seg_network import SegNetwork
    from frtm_vos_b200.model.tracker import Tracker
    from frtm_vos_b200.model.augmenter import ImageAugmenter
    import golden_inputs as GI
    bb = synth.backbone_state_dict(arch, size=size)
    seg = synth.segnet_state_dict(arch)
    C = synth.backbone_out_channels(arch)["layer4"]
    dp = GI.disc_params(C, init_iters=(5, 10, 10, 10) if fast else (5, 10, 10, 10, 10), update_iters=(5,) if fast else (10,),
                        memory_size=memory_size, device=dev)
    fe = ResnetFeatureExtractor(arch, state_dict=bb).to(dev)
    chans = fe.get_out_channels()
    refiner = SegNetwork(1, 64, {L: c for L, c in chans.items() if L != "layer1"}, True)
    trk = Tracker(ImageAugmenter(GI.AUG_PARAMS), fe, dp, refiner, dev)
    trk.load_state_dict(seg)
    trk.to(dev)
    return trk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--objects", type=int, default=3)
    ap.add_argument("--frames", type=int, default=17)
    ap.add_argument("--size", default="480x854")
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    size = tuple(int(v) for v in a.size.split("x"))
    dev = "cuda:0"
    from frtm_vos_b200 import synth, ops
    t0 = time.time()
    trk = build_tracker(a.arch, size, dev, fast=not a.full)
    print("setup %.1fs" % (time.time() - t0))
    seq = synth.SyntheticSequence(num_objects=a.objects, num_frames=a.frames, size=size, seq_id=1)
    seq.preload(dev)
    for rep in range(2):
        l0 = ops.lib().launch_count()
        out, fps = trk.run_sequence(seq)
        print("run %d: %.2f fps, %d launches" % (rep, fps, ops.lib().launch_count() - l0))
    # label quality vs synthetic ground truth
    for t in (1, a.frames // 2, a.frames - 1):
        gt = seq.ground_truth(t)[0]
        lab = out[t].reshape(size).cpu()
        ious = []
        for k in seq.obj_ids:
            inter = ((lab == k) & (gt == k)).sum().item()
            union = ((lab == k) | (gt == k)).sum().item()
            ious.append(inter / max(union, 1))
        print("frame %d IoU vs GT:" % t, ["%.2f" % v for v in ious])
    # stage timings of a tracked frame
    img = seq[a.frames - 1][0]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    def timed(fn, n=5):
        fn(); torch.cuda.synchronize()
        s, e = ev(), ev(); s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    feats, _, _ = trk.feature_extractor.forward_split(img[None])
    print("backbone      %.3f ms" % timed(lambda: trk.feature_extractor.forward_split(img[None])))
    n = a.objects
    scores = torch.randn(n, *feats["layer4"].hi.shape[1:3], device=dev)
    print("seg net (x%d)  %.3f ms" % (n, timed(lambda: trk.refiner.forward_nhwc(scores, feats, size))))
    print("track (all)   %.3f ms" % timed(lambda: trk.track(img)))
    d = trk.targets[1].discriminator
    print("gn update     %.3f ms" % timed(lambda: d.update_optimizer.run(d.update_iters)))
    t0 = time.time(); im5, m5 = trk.augment(img, (seq.ground_truth(a.frames - 1) == 1).byte().to(dev)); torch.cuda.synchronize()
    print("augment       %.1f ms (host)" % ((time.time() - t0) * 1e3))
    _, nh, _ = trk.feature_extractor.forward_split(im5, (), ("layer4",), upto="layer4")
    print("backbone x5   %.3f ms" % timed(lambda: trk.feature_extractor.forward_split(im5, (), ("layer4",), upto="layer4")))
    from frtm_vos_b200.model.discriminator import Discriminator
    import golden_inputs as GI
    def do_init():
        dd = Discriminator(**trk.disc_params)
        dd.init(None, m5, x_nhwc=nh["layer4"])
    print("disc.init     %.3f ms" % timed(do_init, n=2))


if __name__ == "__main__":
    main()
