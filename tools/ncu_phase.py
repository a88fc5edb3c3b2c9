#!/usr/bin/env python
"""Instruction / stall-sample totals per source line from `ncu --page source --csv --print-source cuda,sass`, sorted by
executed warp instructions, plus totals per (file-agnostic) line bucket given as `lo-hi:name` arguments.
usage: python tools/ncu_phase.py x.csv [top] [lo-hi:name ...]"""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
buckets = []
for a in sys.argv[3:]:
    rng, name = a.split(":")
    lo, hi = rng.split("-")
    buckets.append((int(lo), int(hi), name))
rows = list(csv.reader(open(path, newline="")))
hdr = None; cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "Line No":
        hdr = r; isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    if r[0] != "":
        if not r[0].isdigit(): continue
        cur = (int(r[0]), r[1].strip()); agg.setdefault(cur, [0, 0])
        if len(r) <= isamp or r[2] == "": continue
    if cur is None or len(r) <= isamp: continue
    try:
        agg[cur][0] += int(r[isamp] or 0); agg[cur][1] += int(r[iinst] or 0)
    except ValueError:
        pass
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print("samples", ts, "warp instructions", ti)
for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %4d  %s" % (100 * v[1] / ti, 100 * v[0] / ts, ln, src[:120]))
for lo, hi, name in buckets:
    s = sum(v[0] for (ln, _), v in agg.items() if lo <= ln <= hi); i = sum(v[1] for (ln, _), v in agg.items() if lo <= ln <= hi)
    print("bucket %-20s %5.1f%% inst %5.1f%% smp" % (name, 100 * i / ti, 100 * s / ts))
