#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an ncu report captured with --import-source on
(kernel compiled with -lineinfo).  usage: python profiles/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, out, h = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = {c: i for i, c in enumerate(r)}
        ia, ii, isamp = r.index("Address"), r.index("Instructions Executed"), r.index("# Samples")
    elif h and len(r) > ii and r[ia] == "-" and r[0].isdigit():
        try:
            out.append((cur, int(r[0]), r[1].strip()[:110], int(r[ii] or 0), int(r[isamp] or 0)))
        except ValueError:
            pass
tot_i = sum(x[3] for x in out) or 1
tot_s = sum(x[4] for x in out) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("  %inst  %samp  file:line  source")
for f, ln, s, ni, ns in sorted(out, key=lambda x: -x[3])[:top]:
    print(f"  {100 * ni / tot_i:5.1f}  {100 * ns / tot_s:5.1f}  {f}:{ln}  {s}")
