#!/usr/bin/env python
"""profiles/traffic_<workload>.json from an `ncu --set full` report: DRAM bytes per launch of the dominant kernel,
stamped with the hash of the kernel sources it was taken from (bench.py reports `traffic` only while that hash
still matches).   usage: ncu_traffic.py report.ncu-rep kernel-prefix workload [committed-csv-name]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, prefix, workload = sys.argv[1:4]
csv_name = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    if name.replace("void ", "").startswith(prefix):
        rd = float(d["dram__bytes_read.sum"].replace(",", "")); wr = float(d["dram__bytes_write.sum"].replace(",", ""))
        ms = float(d["gpu__time_duration.sum"].replace(",", "")) / 1e6
        if best is None or rd > best["dram_bytes_read"]:
            best = {"kernel": name.replace("void ", "")[:80], "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "duration_ms_under_ncu": ms}
assert best, "kernel not found"
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
best.update({"workload": workload, "csrc_sha": b.csrc_hash(),
             "source": f"{csv_name} (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)",
             "csrc_sha_covers": list(b.STREAM_KERNEL_SOURCES)})
path = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
json.dump(best, open(path, "w"), indent=1)
print(path, best)
