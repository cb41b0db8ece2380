#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys, collections, subprocess
rep = sys.argv[1]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = []
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit():
        try:
            n = int(r[7] or 0); st = int(r[4] or 0)
        except ValueError:
            continue
        agg.append((n, st, cur_file, int(r[0]), r[1].strip()[:110])); tot += n
agg.sort(reverse=True)
print(f"total warp-instr {tot:.4g} -> {tot*32/div:.1f} per unit")
for n, st, f, ln, src in agg[:top]:
    print(f"{n*32/div:7.1f} {100*n/tot:5.1f}% stall {st:7d}  {f}:{ln}  {src}")
