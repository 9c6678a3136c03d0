#!/usr/bin/env python
"""Mnemonic counts per kernel from `cuobjdump -sass` of the built library -> profiles/r2_sass_summary.txt
(the .so itself is git-ignored; this is the committed record that the kernels are tcgen05 / TMEM / TMA code).
    python tools/sass_summary.py [path/to/libmclip_b200.so] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mamba_clip_b200", "libmclip_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "HMMA", "FFMA",
         "RED", "ATOM", "UBLKCP"]
cur = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    c = counts[cur]
    c["total"] += 1
    two_cta = op.startswith("UTCHMMA") and ".2CTA" in op
    for w in WATCH:
        if w == "UTCHMMA.2CTA":
            c[w] += int(two_cta)
        elif w == "UTCHMMA":
            c[w] += int(op.startswith("UTCHMMA") and not two_cta)
        elif op == w or op.startswith(w + "."):
            c[w] += 1


def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        return d.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    except Exception:
        return n


print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction mnemonic counts per kernel (sm_100a)")
print(f"# columns: total | " + " | ".join(WATCH))
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k).replace("mclip::", "")
    print(f"{name[:70]:70s} {c['total']:7d} | " + " | ".join(f"{c[w]:5d}" for w in WATCH))
    tot.update(c)
print(f"{'ALL KERNELS':70s} {tot['total']:7d} | " + " | ".join(f"{tot[w]:5d}" for w in WATCH))
