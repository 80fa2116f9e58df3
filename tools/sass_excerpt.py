#!/usr/bin/env python
"""profiles/sass_excerpt.txt: which Blackwell instructions the built library really contains, per kernel (cuobjdump -sass of
leaf_b200/lib/libleaf_b200.so): tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG / UBLKCP), mbarrier (SYNCS), 256-bit global
accesses (LDG.E.ENL2.256 / STG.E.ENL2.256), warp-level tensor ops (HMMA), plus the full instruction lines of the GEMM's MMA issue
loop. No GPU needed.   python tools/sass_excerpt.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leaf_b200", "lib", "libleaf_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "ELECT", "HMMA", "LDSM", "MOVM", "LDG.E.ENL2.256",
        "STG.E.ENL2.256", "LDGSTS", "MUFU", "RED", "ATOM"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fn, per, lines = None, collections.OrderedDict(), collections.defaultdict(list)
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[fn] = collections.Counter()
        continue
    if fn is None or "/*" not in ln:
        continue
    body = ln.split("*/", 1)[1] if "*/" in ln else ln
    per[fn]["instructions"] += 1 if re.search(r"^\s+[A-Z@!]", body) else 0
    for k in KEYS:
        if re.search(r"\b" + re.escape(k), body):
            per[fn][k] += 1
            if k in ("UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "UBLKCP") and len(lines[(fn, k)]) < 6:
                lines[(fn, k)].append(body.strip().rstrip(";").split("/*")[0].strip())
out = [f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)", ""]
for fn, c in per.items():
    hits = {k: v for k, v in c.items() if k != "instructions" and v}
    if hits:
        out.append(f"{fn}: {c['instructions']} SASS instructions  " + "  ".join(f"{k} x{v}" for k, v in hits.items()))
out.append("")
for (fn, k), ls in lines.items():
    out.append(f"--- {fn} :: {k}")
    out += ["    " + l for l in ls]
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_excerpt.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
