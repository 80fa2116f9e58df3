#!/usr/bin/env python
"""profiles/gemm_traffic.json from ncu --set full reports of GEMM launches (dram__bytes_read.sum + dram__bytes_write.sum
per launch, averaged over the captured launches of both phases): python tools/ncu_traffic.py rep1.ncu-rep rep2.ncu-rep"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot, n, detail = 0.0, 0, []
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        tot += b
        n += 1
        detail.append(dict(kernel=r[ik][:40], dram_bytes=b, duration=r[it] + " " + units[it]))
json.dump(dict(workload="ViT-H-14 text tower, LEAF k=1 rho=50, batch 128 typical synthetic captions per GPU",
               dram_bytes_per_launch=tot / n, launches=n, detail=detail,
               how="ncu --set full --clock-control none, 4 GEMM launches (one layer) of each attack phase"),
          open("profiles/gemm_traffic.json", "w"), indent=1)
print(tot / n, n)
