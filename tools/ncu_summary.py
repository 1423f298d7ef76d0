#!/usr/bin/env python
"""One kernel of an .ncu-rep -> the selected-rows CSV committed under profiles/ (metric, unit, value).
usage: ncu_summary.py report.ncu-rep kernel-regex out.csv [label]"""
import csv, io, re, subprocess, sys
rep, pat, out = sys.argv[1:4]
label = sys.argv[4] if len(sys.argv) > 4 else pat
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
keep = re.compile(r"^(Kernel Name|dram__|gpu__dram|gpu__time_duration|launch__|sm__warps_active|sm__inst_executed_pipe_(fp64|alu|fma|lsu)|"
                  r"smsp__issue_active|smsp__inst_executed\.sum$|smsp__thread_inst_executed_per_inst|smsp__average_warps?_issue_stalled_.*per_warp_active|"
                  r"smsp__average_warp_latency|l1tex__t_bytes|lts__t_bytes|sm__throughput|smsp__warp_issue_stalled.*\.pct$)")
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if re.search(pat, d["Kernel Name"]):
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        if best is None or t > best[0]:
            best = (t, r)
assert best, "kernel not found"
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", label])
    for h, u, v in zip(hdr, units, best[1]):
        if keep.match(h):
            w.writerow([h, u, v])
print(out, "duration", best[0])
