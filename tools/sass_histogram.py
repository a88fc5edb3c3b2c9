#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of the in-tree libcopra_b200.so (evidence for profiles/: DMMA / TMA bulk copy /
128-bit loads / redux / cluster barriers).  usage: cuobjdump -sass copra_b200/lib/libcopra_b200.so | python tools/sass_histogram.py"""
import collections
import re
import sys

txt = sys.stdin.read()
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "LDG.E.128.CONSTANT", "LDG.E.128", "LDG.E.64.CONSTANT", "LDG.E.64", "STG.E.128", "LDS.128", "LDS.64",
        "STS.128", "SHFL.BFLY", "REDUX", "UTMALDG", "UBLKCP.S.G", "SYNCS", "LDGSTS", "UCGABAR", "BAR.SYNC", "ATOMG", "MUFU.RSQ64H", "MUFU.RCP64H"]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z][A-Za-z0-9_.]+)", f, re.M)
    cnt = collections.Counter(ops)
    row = []
    used = set()
    for k in KEYS:
        v = sum(c for o, c in cnt.items() if o.startswith(k) and o not in used)
        used.update(o for o in cnt if o.startswith(k))
        if v:
            row.append("%s=%d" % (k, v))
    print("%-90s total=%-6d %s" % (name[:90], len(ops), " ".join(row)))
