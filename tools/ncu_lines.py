#!/usr/bin/env python
"""Per-CUDA-source-line summary of one kernel from an ncu report captured with --import-source on (built with -lineinfo):
   python tools/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ie, ss, wx = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared Excessive")
        sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and r[0].isdigit() and len(r) > ie:
        try:
            n, s, w = int(r[ie] or 0), int(r[ss] or 0), int(r[wx] or 0)
        except ValueError:
            continue
        st = sorted(((hdr[i][6:], int(r[i])) for i in sc if r[i] not in ("", "0", "-")), key=lambda kv: -kv[1])[:2]
        out.append((s, n, w, cur_file, int(r[0]), r[1].strip()[:90], st))
ti, ts = sum(o[1] for o in out), sum(o[0] for o in out)
print(f"total warp-instructions {ti}, samples {ts}")
for s, n, w, f, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{f[:12]:12s}:{ln:4d} inst {100 * n / ti:5.1f}% samp {100 * s / ts:5.1f}% xwf {w:8d} | {src:90s} | {st}")
