#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into a markdown table for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time (us, under ncu)"),
    ("dram__bytes_read.sum", "DRAM read (MB)"),
    ("dram__bytes_write.sum", "DRAM written (MB)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1TEX data pipe % of peak"),
    ("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "L1TEX data-pipe wavefronts / SM"),
    ("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "  of which global"),
    ("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "  of which shared/shuffle"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "static smem / block (B)"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr = r[0]
    return hdr, r[1], r[2:]


print("| report | kernel | " + " | ".join(k[1] for k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    hdr, units, rows = rows_of(rep)
    kn = hdr.index("Kernel Name")
    for r in rows:
        vals = []
        for k, _ in KEYS:
            v = r[hdr.index(k)] if k in hdr else ""
            try:
                f = float(v)
                u = units[hdr.index(k)] if k in hdr else ""
                # ncu picks a unit per column: normalise times to us and byte counts to MB (the column titles say so)
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if k == "launch__shared_mem_per_block_static":
                    f = float(v) * {"byte": 1.0, "Kbyte": 1e3}.get(u, 1.0)
                v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(v)
        name = r[kn].split("(")[0].replace("void ", "")[:48]
        print(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(vals) + " |")
