#!/usr/bin/env python3
"""Per-source-line instruction and stall shares of one kernel of an `ncu --set full --import-source on` report.
usage: ncu_source_hotspots.py <report.ncu-rep> <kernel regex> [min share, default 0.006] > profiles/<name>.txt
Columns: share of the kernel's executed warp instructions, share of its stall samples, threads per instruction,
SASS instructions attributed to the line (all template instantiations together), source line."""
import re
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.006
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kernel}"], capture_output=True, text=True).stdout
# line rows: "<line>","<source>","-","-","<samples all>","<samples not issued>","<# samples>","<inst>","<thread inst>",...
pat = re.compile(r'^"(\d+)","(.*?)","-","-","(\d+)","(\d+)","(\d+)","(\d+)","(\d+)"')
groups = []
for line in raw.splitlines():
    m = pat.match(line)
    if m:
        groups.append([int(m.group(1)), m.group(2), float(m.group(6)), float(m.group(7)), float(m.group(5)), 0])
    elif line.startswith('"","","0x') and groups:
        groups[-1][5] += 1
ti, ts = sum(g[2] for g in groups), sum(g[4] for g in groups)
print(f"# {kernel}: {ti:.0f} warp instructions, {sum(g[3] for g in groups) / max(ti, 1):.1f} threads per instruction, {ts:.0f} stall samples")
print("# line   %inst  %stall  thr/inst  sass  source")
for ln, src, ins, thr_i, smp, n in groups:
    if ins / ti > thr or smp / ts > thr * 1.3:
        print(f"{ln:6d}  {100 * ins / ti:5.2f}  {100 * smp / ts:6.2f}  {thr_i / max(ins, 1):8.1f}  {n:4d}  {src.strip()[:110]}")
