"""Per-launch summary of an ncu --set full report: python tools/ncu_summary.py <report.ncu-rep> "<comment>" > profiles/x.csv
(reads `ncu -i <report> --page raw --csv`; one row per captured launch)."""
import csv, io, subprocess, sys

rep, note = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = [("time_us", "gpu__time_duration.sum", 1.0), ("dram_read_bytes", "dram__bytes_read.sum", 1.0),
        ("dram_write_bytes", "dram__bytes_write.sum", 1.0),
        ("tensor_pipe_active_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("tensor_mem_active_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("dram_throughput_pct", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1.0),
        ("registers", "launch__registers_per_thread", 1.0), ("sm_clock_mhz", "sm__cycles_elapsed.avg.per_second", 1.0)]
unit_scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6,
              "Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6, "cycle/nsecond": 1e3, "cycle/usecond": 1.0, "cycle/second": 1e-6}
units = rows[1]
ki = hdr.index("Kernel Name")
out = csv.writer(sys.stdout)
if note:
    out.writerow(["# " + note])
out.writerow(["kernel"] + [c[0] for c in cols])
for r in rows[2:]:
    if len(r) <= ki:
        continue
    name = r[ki].replace("void ", "").replace("cbas::", "").replace("<unnamed>::", "")
    name = name.split("(")[0]
    vals = []
    for _, metric, _s in cols:
        if metric not in hdr:
            vals.append("")
            continue
        i = hdr.index(metric)
        try:
            v = float(r[i].replace(",", "")) * unit_scale.get(units[i], 1.0)
            vals.append(f"{v:.6g}")
        except ValueError:
            vals.append("")
    out.writerow([name] + vals)
