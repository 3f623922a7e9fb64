#!/usr/bin/env python
"""SASS opcode summary of the shipped library (no GPU needed): per kernel, how many instructions of the kinds that
prove a design claim - bulk-async / tensor TMA copies (UBLKCP / UTMALDG), mbarrier traffic (SYNCS), shared / global
atomics (ATOMS / ATOMG / RED), the native 16x2 three-input min/max (VIMNMX3), XU-pipe conversions (I2F / F2I / MUFU).
usage: python profiles/sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lfd_b200", "liblfd_b200.so")
OPS = ["UBLKCP", "UTMALDG", "SYNCS", "ATOMS", "ATOMG", "RED", "VIMNMX3", "VIMNMX", "I2F", "F2I", "MUFU", "SHFL", "VOTE", "LDS", "STS",
       "LDG", "STG", "BAR", "FADD", "FMUL", "FFMA"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
cur = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_all"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
                break
        else:
            if op.startswith("VIMNMX3"):
                cur["VIMNMX3"] += 1
print("# SASS opcode summary of %s\n" % os.path.relpath(lib, ROOT))
print("Architectures in the cubin: %s.  Counts are static instructions per kernel (`cuobjdump -sass`), not executed ones.\n" % ", ".join(arch))
cols = ["_all"] + OPS
print("| kernel | " + " | ".join("total" if c == "_all" else c for c in cols) + " |")
print("|---|" + "---:|" * len(cols))
for name, c in sorted(counts.items(), key=lambda kv: -kv[1]["_all"]):
    if c["_all"] < 40:
        continue
    print("| %s | " % name + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |")
