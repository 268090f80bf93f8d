#!/usr/bin/env python
"""Per-source-line hot spots of one kernel launch in an ncu report (read on the build container, no GPU needed):
`ncu -i <report> --page source --csv --print-source cuda,sass`, the SASS rows summed up under the CUDA line they were compiled
from.  Columns: warp-stall samples (share of the launch's samples, and the three largest stall reasons), instructions executed,
L1 tag requests (global), shared-memory wavefronts, L2 sectors (global).
usage: python scripts/ncu_source_hotspots.py profiles/<report>.ncu-rep <launch index> [top N] >> profiles/<out>.md"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

rep, launch = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

kernel, files, cur_file, hdr, cur_line = None, {}, None, None, None
acc = defaultdict(lambda: defaultdict(float))
want = {"samples": "Warp Stall Sampling (All Samples)", "inst": "Instructions Executed", "l1tag": "L1 Tag Requests Global",
        "smem": "L1 Wavefronts Shared", "l2": "L2 Theoretical Sectors Global"}
stalls = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        stalls = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur_line = (cur_file, int(r[0]))
        continue
    if len(r) > 2 and r[0] == "" and r[2].startswith("0x") and len(r) == len(hdr) and cur_line:
        def num(i):
            try:
                return float(r[i])
            except ValueError:
                return 0.0
        for k, name in want.items():
            if name in hdr:                      # (kernels without shared memory have no shared-memory columns)
                acc[cur_line][k] += num(hdr.index(name))
            else:
                acc[cur_line][k] += 0.0
        for i, n in stalls:
            acc[cur_line][n] += num(i)
        acc[cur_line]["nsass"] += 1

tot = {k: sum(v[k] for v in acc.values()) for k in want}


def src(file, line):
    p = file if os.path.exists(file) else os.path.join(ROOT, file.split("/root/repo/")[-1])
    try:
        return open(p).read().splitlines()[line - 1].strip()[:110]
    except Exception:
        return "?"


print(f"### `{kernel}` — launch {launch} of `{os.path.basename(rep)}`\n")
print(f"Totals over the kernel: {int(tot['samples'])} warp-stall samples, {int(tot['inst'])} warp instructions executed, "
      f"{int(tot['l1tag'])} L1 tag requests (global), {int(tot['smem'])} shared-memory wavefronts, {int(tot['l2'])} L2 sectors (global).\n")
print("| source line | stall samples (share) | largest stall reasons | instr. executed (share) | L1 tag req. global (share) | smem wavefronts | L2 sectors global (share) |")
print("|---|---|---|---|---|---|---|")
for (f, ln), v in sorted(acc.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    rs = sorted(((v[n], n) for _, n in stalls), reverse=True)[:3]
    rs = ", ".join(f"{n.replace('stall_', '')} {100 * x / max(v['samples'], 1):.0f} %" for x, n in rs if x > 0)
    sh = lambda k: f"{int(v[k])} ({100 * v[k] / max(tot[k], 1):.1f} %)"
    print(f"| `{os.path.basename(f)}:{ln}` `{src(f, ln).replace('|', '¦')}` | {sh('samples')} | {rs} | {sh('inst')} | {sh('l1tag')} | {int(v['smem'])} | {sh('l2')} |")
print()
