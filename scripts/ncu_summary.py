#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep (ncu --set full --import-source on): key raw metrics, stall reasons per issued
instruction, dynamic opcode mix per work unit, top stall sites.   usage: ncu_summary.py REPORT UNITS_PER_LAUNCH [title]"""
import collections
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, unit, val = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
print(f"# {title}")
print(f"# kernel: {m['Kernel Name'][0][:110]}   grid {m.get('Grid Size', ('?',))[0]}  block {m.get('Block Size', ('?',))[0]}")
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in m:
        print(f"{k:85s} {m[k][0]:>16s} {m[k][1]}")
try:
    wi = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
    print(f"warp instructions per unit x32 = {wi * 32 / units:.1f} thread-instructions per unit ({units:.4g} units per launch)")
    db = sum(float(m[k][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(f"DRAM bytes per unit = {db / units:.1f}")
except Exception as e:
    print("derived metrics unavailable:", e)
print("stall cycles per issued instruction:")
st = [(float(v[0]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, v in m.items() if "issue_stalled_" in h and "_per_issue_active.ratio" in h]
for v, h in sorted(st, reverse=True)[:9]:
    print(f"   {h:22s} {v:6.2f}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
sh = src[1]
iS, iN, iSt = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
mix, stall = collections.Counter(), collections.Counter()
tot = tots = 0
rows = []
for r in src[2:]:
    if len(r) <= iN:
        continue
    t = r[iS].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0] if not op.startswith("IMAD") else ("IMAD.MOV" if "MOV" in op else ("IMAD.WIDE" if "WIDE" in op else "IMAD"))
    n, s = int(r[iN] or 0), int(r[iSt] or 0)
    mix[op] += n
    stall[op] += s
    tot += n
    tots += s
    rows.append((s, n, r[iS].strip()[:90]))
print(f"dynamic opcode mix (thread-instructions per unit, total {tot * 32 / units:.1f}):")
for op, n in mix.most_common(22):
    print(f"   {op:12s} {n * 32 / units:7.1f}  {100 * n / tot:5.1f}%   stall samples {100 * stall[op] / max(tots, 1):5.1f}%")
print("top stall sites (share of all samples, executed count, SASS):")
for s, n, t in sorted(rows, reverse=True)[:14]:
    print(f"   {100 * s / max(tots, 1):5.1f}%  {n:12d}  {t}")
