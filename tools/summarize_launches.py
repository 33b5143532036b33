"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/summarize_launches.py file.csv [--md title]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((name, v))
    return agg, order


if __name__ == "__main__":
    agg, order = load(sys.argv[1])
    tot = sum(a[1] for a in agg.values())
    md = "--md" in sys.argv
    if md:
        print("total kernel time %.3f ms over %d launches\n" % (tot / 1e3, sum(a[0] for a in agg.values())))
        print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    else:
        print("total us %.1f launches %d" % (tot, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, n, t, 100 * t / tot, t / n))
        else:
            print("%-58s n=%5d %10.1f us %5.1f%% avg %8.1f" % (k[:58], n, t, 100 * t / tot, t / n))
    if "--seq" in sys.argv:
        for name, v in order:
            print("%8.1f  %s" % (v, name[:70]))
