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
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("| metric | value |\n|---|---|")
    print("| kernel | `%s` |" % d.get("Kernel Name", "?")[:90])
    for k in WANT:
        if k in d:
            print("| %s | %s %s |" % (k, d[k], u.get(k, "")))
    print()
