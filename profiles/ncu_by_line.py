#!/usr/bin/env python
"""Dynamic instructions + stall samples per CUDA source line from an ncu report (needs --import-source on / -lineinfo).
usage: ncu_by_line.py rep.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r
        iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
        try:
            ex, smp = int(r[iex] or 0), int(r[ismp] or 0)
        except ValueError:
            continue
        st = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:3]
        recs.append((fname, int(r[0]), ex, smp, r[1].strip()[:70], st))
tex, tsm = sum(r[2] for r in recs), sum(r[3] for r in recs)
print(f"total inst {tex}  samples {tsm}")
for f, ln, ex, smp, src, st in recs:
    if ex * 100.0 / max(tex, 1) >= minpct or smp * 100.0 / max(tsm, 1) >= minpct:
        print(f"{f:22s}{ln:5d} inst {ex*100.0/tex:5.1f}%  smp {smp*100.0/max(tsm,1):5.1f}%  {src:70s} " + " ".join(f"{h}:{n}" for n, h in st if n))
