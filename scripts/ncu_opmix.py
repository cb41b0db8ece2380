#!/usr/bin/env python
"""Dynamic opcode mix of a kernel from an `ncu --page source --csv` dump: warp-level executed instructions per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iN: continue
    src = r[iS].split()
    if not src: continue
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0] if not op.startswith("IMAD") else ("IMAD.MOV" if "MOV" in op else ("IMAD.WIDE" if "WIDE" in op else "IMAD"))
    n = int(r[iN]); mix[op] += n; tot += n; stall[op] += int(r[iSt] or 0)
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total warp-instr {tot:.4g}; per unit x32/{div:.4g} = {tot*32/div:.1f}")
for op, n in mix.most_common(28):
    print(f"{op:12s} {n:14d} {100*n/tot:5.1f}%  per-unit {n*32/div:7.1f}  stall-samples {stall[op]}")
