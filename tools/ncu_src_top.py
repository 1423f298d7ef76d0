#!/usr/bin/env python
"""Top source lines of one kernel of an .ncu-rep by executed warp instructions and stall samples.
usage: ncu_src_top.py report.ncu-rep kernel-name [N]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
fname = "?"; hdr = None
agg = collections.OrderedDict()
for row in csv.reader(io.StringIO(out)):
    if len(row) >= 2 and row[0] == "File Name": fname = row[1].split("/")[-1]; continue
    if len(row) > 4 and row[0] == "Line No": hdr = row; continue
    if hdr and len(row) == len(hdr) and row[0] != "":
        d = dict(zip(hdr, row))
        try: inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
        except ValueError: continue
        src = row[1].strip()
        k = (fname, int(row[0]))
        a = agg.setdefault(k, [0, 0, src, collections.Counter()])
        a[0] += inst; a[1] += samp
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                try: a[3][h] += int(d[h])
                except ValueError: pass
tot_i = sum(a[0] for a in agg.values()) or 1; tot_s = sum(a[1] for a in agg.values()) or 1
byfile = collections.Counter(); byfile_s = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[0]; byfile_s[f] += a[1]
print("total warp instr", tot_i, "samples", tot_s)
for f, v in byfile.most_common(): print(f"  {f:28s} instr {100*v/tot_i:5.1f}%  samples {100*byfile_s[f]/tot_s:5.1f}%")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in a[3].most_common(3))
    print(f"{f}:{l:4d} instr {100*a[0]/tot_i:5.1f}% samp {100*a[1]/tot_s:5.1f}%  [{top}]  {a[2][:90]}")
