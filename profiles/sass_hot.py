#!/usr/bin/env python
"""Per-region instruction/stall breakdown from an ncu report's SASS page.
usage: sass_hot.py rep.ncu-rep [kernel-index]  -> prints 64-instruction buckets with executed counts and samples"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("total inst executed", tot, "samples", tots, "n_sass", len(data))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for i in range(0, len(data), B):
    chunk = data[i:i + B]
    e = sum(d[1] for d in chunk); s = sum(d[2] for d in chunk)
    if e * 200 > tot or s * 200 > tots:
        fp64 = sum(d[1] for d in chunk if d[0].split()[0].lstrip('@!P0123456789 ').startswith(('DFMA', 'DADD', 'DMUL')) or (len(d[0].split()) > 1 and d[0].split()[1].startswith(('DFMA', 'DADD', 'DMUL'))))
        print(f"[{i:5d}-{i+len(chunk):5d}) exec {e/tot*100:5.1f}%  samples {s/tots*100:5.1f}%  fp64 {fp64/max(e,1)*100:4.0f}%  first: {chunk[0][0][:50]}")
