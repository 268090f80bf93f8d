#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of the dominant kernel (read here, no GPU needed):
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed to the hash of the kernel sources it was captured from
(bench.py quotes it as roofline.traffic only while the sources are unchanged).
usage: python scripts/make_traffic.py profiles/<report>.ncu-rep "b200::csr_flat_kernel<double>" """
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
short = kernel.split("::")[-1].split("<")[0]
r = [x for x in rows[2:] if short in x[hdr.index("Kernel Name")]][0]
rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
unit = rows[1][hdr.index("dram__bytes_read.sum")]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
unit_w = rows[1][hdr.index("dram__bytes_write.sum")]
scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit_w]
rec = {"source_sha": kernel_source_sha(kernel), "source_sha_of": "the .cu file defining the kernel + spmv_common.cuh + nvcc flags (bench.py kernel_source_sha)", "kernel": kernel, "workload": "R-MAT 1M x 1M, 16M nnz, fp64 (BASELINE.json configs[1])",
       "dram_bytes_per_launch": int(rd * scale + wr * scale_w), "dram_read_bytes": int(rd * scale), "dram_write_bytes": int(wr * scale_w),
       "kernel_time_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")]), "ncu_report": os.path.relpath(rep, ROOT)}
json.dump(rec, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(rec)
