#!/usr/bin/env python
"""Warp instructions / stall samples of an ncu report (--import-source on, -lineinfo) summed per `// @region name`
block of the CUDA sources.  usage: python profiles/ncu_regions.py rep.ncu-rep units_per_launch [unit_name]"""
import bisect, csv, io, os, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
uname = sys.argv[3] if len(sys.argv) > 3 else "unit"
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2fast2q_b200", "csrc")
regions = {}
for f in os.listdir(CSRC):
    marks = [(1, "(top)")]
    for i, line in enumerate(open(os.path.join(CSRC, f)), 1):
        if "// @region" in line:
            marks.append((i, line.split("@region")[1].strip()))
    regions[f] = marks
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h, acc = None, None, {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ia, ii, isamp, ith = r.index("Address"), r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
        h = True
    elif h and len(r) > ii and r[ia] == "-" and r[0].isdigit():
        ln = int(r[0])
        if cur in regions:
            m = regions[cur]
            name = f"{cur}:{m[bisect.bisect_right([x[0] for x in m], ln) - 1][1]}"
        else:
            name = cur
        a = acc.setdefault(name, [0, 0, 0])
        try:
            a[0] += int(r[ii] or 0); a[1] += int(r[isamp] or 0); a[2] += int(r[ith] or 0)
        except ValueError:
            pass
ti, ts = sum(a[0] for a in acc.values()) or 1, sum(a[1] for a in acc.values()) or 1
print(f"total warp instructions {ti} ({ti / units:.1f} per {uname}), samples {ts}")
print(f"  {'region':44s} %inst  %samp  warp-inst/{uname}  lanes")
for k, a in sorted(acc.items(), key=lambda x: -x[1][0]):
    if a[0] * 500 > ti:
        print(f"  {k:44s} {100 * a[0] / ti:5.1f}  {100 * a[1] / ts:5.1f}  {a[0] / units:8.1f}  {a[2] / max(a[0], 1):5.1f}")
