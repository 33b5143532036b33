"""Key metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`) as markdown tables:
python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
    # python tools/ncu_summary.py report.ncu-rep --traffic "<workload name>" <active samples> <kernel regex> <out.json>
    RUN_TRAFFIC = True
else:
    RUN_TRAFFIC = False
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout if not RUN_TRAFFIC else ""
rows = list(csv.reader(io.StringIO(out))) if not RUN_TRAFFIC else [[], []]
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("| metric | value |\n|---|---|")
    print("| kernel | `%s` |" % d.get("Kernel Name", "?")[:90])
    for k in WANT:
        if k in d:
            print("| %s | %s %s |" % (k, d[k], u.get(k, "")))
    print()


def traffic_json(rep, workload, active_samples, kernel_regex, out_path):
    """Adds {workload: {dram_bytes_per_launch, active_samples, kernel, source}} to the JSON bench.py reads for roofline.traffic."""
    import json, os, re
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    vals = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, rows[1]))
        if not re.search(kernel_regex, d.get("Kernel Name", "")):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(d[k].replace(",", ""))
            tot += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
        vals.append((tot, d["Kernel Name"]))
    assert vals, "no launch of %s in %s" % (kernel_regex, rep)
    data = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    data[workload] = dict(dram_bytes_per_launch=sum(v for v, _ in vals) / len(vals), active_samples=int(active_samples),
                          kernel=vals[0][1].split("(")[0], launches_captured=len(vals), source=os.path.basename(rep))
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data[workload]))


if RUN_TRAFFIC:
    traffic_json(sys.argv[1], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6])
