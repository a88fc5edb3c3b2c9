#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="")))
fname, hdr, cur = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_idx = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        iexc = hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if hdr is None:
        continue
    if r[0] != "":
        if not r[0].isdigit():
            continue
        cur = (fname, int(r[0]), r[1].strip())
        agg.setdefault(cur, dict(samples=0, inst=0, exc=0, stalls={}))
        if len(r) <= isamp or r[2] == "":
            continue
    if cur is None or len(r) <= isamp:
        continue
    a = agg[cur]
    try:
        a["samples"] += int(r[isamp] or 0)
        a["inst"] += int(r[iinst] or 0)
        a["exc"] += int(r[iexc] or 0)
        for i, h in stall_idx:
            v = int(r[i] or 0)
            if v:
                a["stalls"][h] = a["stalls"].get(h, 0) + v
    except ValueError:
        pass
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join("%s=%d%%" % (k[6:], 100 * v / max(1, a["samples"])) for k, v in st)
    print("%5.1f%% smp %5.1f%% inst exc=%-8d %s:%d  %s   [%s]" % (100 * a["samples"] / tot, 100 * a["inst"] / toti, a["exc"], f, ln, src[:90], sts))
