#!/usr/bin/env python
"""Per-source-line stall-sample summary of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/srcprof.py gpurun_out/prof.ncu-rep nms_image [top_n]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, acc = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":      # a CUDA source line row (aggregated over its SASS)
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        key = (fname, int(r[0]))
        s, n = float(r[si] or 0), float(r[ii] or 0)
        a = acc.setdefault(key, [0.0, 0.0, r[1]])
        a[0] += s
        a[1] += n
tot = sum(a[0] for a in acc.values()) or 1.0
print(f"kernel ~{kern}: {tot:.0f} stall samples")
for (f, ln), (s, n, src) in sorted(sorted(acc.items(), key=lambda kv: -kv[1][0])[:top_n]):
    print(f"{f}:{ln:<4d} {100 * s / tot:5.1f}%  inst={n:>9.0f}  {src.strip()[:100]}")
