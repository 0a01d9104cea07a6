#!/usr/bin/env python
"""Per-instruction view of an ncu report: executed count, samples and the dominant stall reasons.
usage: sass_lines.py rep.ncu-rep first last   (instruction index range)"""
import csv, io, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(rows[2:]):
    if a <= k < b:
        st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
        print(f"{k:5d} {int(r[iex] or 0):9d} {int(r[ismp] or 0):5d}  {r[isrc].strip()[:60]:60s} " + " ".join(f"{h}:{n}" for n, h in st if n))
