"""Reference-vs-reference spread of free-running label maps (finding 8): the CPU oracle (bit-identical to the
executed reference on the golden inputs) is run on the P5 test sequence with different thread counts — identical inputs,
only the reduction order inside ATen changes — and the label agreement between the runs is reported.  This is the noise
floor any free-running comparison (tests/test_gpu_model.py::test_P5_end_to_end_free_running) sits on.

    python tools/oracle_spread.py [threads ...]        (default 1 4 8)
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import golden_inputs as GI  # noqa: E402
from test_gpu_model import _e2e_setup, _oracle_tracker  # noqa: E402

threads = [int(v) for v in sys.argv[1:]] or [1, 4, 8]
bb, seg, dp, seq, size = _e2e_setup()
runs = {}
for n in threads:
    torch.set_num_threads(n)
    orc = _oracle_tracker(bb, seg, dp)
    torch.manual_seed(11)
    out, _ = orc.run_sequence(seq)
    runs[n] = torch.stack([o.reshape(size) for o in out])
    print("threads=%d done" % n, flush=True)
res = {}
for i, a in enumerate(threads):
    for b in threads[i + 1:]:
        agree = (runs[a] == runs[b]).float().mean().item()
        worst = min((runs[a][t] == runs[b][t]).float().mean().item() for t in range(runs[a].shape[0]))
        res["%d_vs_%d" % (a, b)] = dict(agreement=agree, worst_frame=worst, differing_px=int((runs[a] != runs[b]).sum()))
print(json.dumps(dict(sequence="2 objects, 18 frames, %dx%d (the P5 test case)" % size, pairs=res), indent=1))
