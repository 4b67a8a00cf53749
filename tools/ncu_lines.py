#!/usr/bin/env python
"""Top source lines of an ncu report by executed instructions / stall samples.
usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        try:
            ie = int(d["Instructions Executed"]); sm = int(d["# Samples"])
        except Exception: continue
        lines.append((ie, sm, cur_file, r[0], r[1].strip()[:110], d.get("L1 Wavefronts Shared Excessive", "0")))
tot_i = sum(l[0] for l in lines); tot_s = sum(l[1] for l in lines)
print("total instr %d samples %d" % (tot_i, tot_s))
key = (lambda x: -x[1]) if (len(sys.argv) > 3 and sys.argv[3] == "samp") else (lambda x: -x[0])
for l in sorted(lines, key=key)[:top]:
    print("%5.1f%% inst %5.1f%% samp  %s:%s  %s   [smem excess wf %s]" % (100.0*l[0]/max(tot_i,1), 100.0*l[1]/max(tot_s,1), l[2], l[3], l[4], l[5]))
