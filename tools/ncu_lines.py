"""Warp-stall samples of an .ncu-rep aggregated per CUDA source line (needs -lineinfo and --import-source on):
python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv, collections, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, ix, stalls = None, None, {}, []
agg = collections.OrderedDict()
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr, ix = r, {}
        for i, n in enumerate(hdr):
            ix.setdefault(n, i)
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue
    ns = int(r[ix["# Samples"]]) if r[ix["# Samples"]].strip().isdigit() else 0
    if ns == 0:
        continue
    key = (cur, int(r[0]))
    st = {n: (int(r[ix[n]]) if r[ix[n]].strip().isdigit() else 0) for n in stalls}
    if key in agg:
        ns0, src, st0 = agg[key]
        agg[key] = (ns0 + ns, src, {k: st0[k] + st[k] for k in st})
    else:
        agg[key] = (ns, r[1].strip()[:100], st)
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
by_stall = collections.Counter()
for ns, src, st in agg.values():
    by_stall.update(st)
print("by reason:", " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(sum(by_stall.values()), 1)) for k, v in by_stall.most_common(8)))
for (f, ln), (ns, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %s:%d  %s   [%s]" % (100 * ns / tot, f, ln, src, " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))
