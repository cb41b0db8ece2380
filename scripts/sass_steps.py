#!/usr/bin/env python
"""Static per-row-step instruction mix of a marching kernel from `cuobjdump -sass` output: splits the listing at BAR.SYNC and
prints the opcode histogram of the window [bar k0, bar k1) -- with k1 - k0 = 3 that is one row step of k_advect5."""
import collections
import re
import sys

ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append(m.group(2).strip())
bars = [k for k, t in enumerate(ins) if 'BAR.SYNC' in t]
print("instructions", len(ins), "barriers", len(bars))
print("gaps", [b - a for a, b in zip(bars, bars[1:])])
if len(sys.argv) > 3:
    k0, k1 = int(sys.argv[2]), int(sys.argv[3])
    seg = ins[bars[k0]:bars[k1]]
    c = collections.Counter()
    for t in seg:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = t.split()[0]
        c[op if op.startswith('IMAD') else op.split('.')[0]] += 1
    print(len(seg), c.most_common())
    if len(sys.argv) > 4:
        print("\n".join(seg))
