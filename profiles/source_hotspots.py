#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per CUDA source line.
usage: ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:<kernel> > k.csv
       python profiles/source_hotspots.py k.csv [top_n]
Prints, per source line, its share of the kernel's warp-stall samples and of its executed warp instructions
(first launch of the kernel in the report; SASS rows shared by several source lines are counted once)."""
import csv, sys, collections
path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
cur_file = None; cur_line = None; cur_src = None
agg = collections.OrderedDict(); seen = set()
h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; ci = h.index("# Samples"); ii = h.index("Instructions Executed"); continue
    if h is None or len(r) <= ci: continue
    if r[0] != "":
        cur_line = (cur_file, r[0]); cur_src = r[1]
        agg.setdefault(cur_line, [cur_src, 0, 0])
    elif r[2] not in ("...", "-", ""):
        key = (cur_line, r[2])
        if key in seen: continue
        seen.add(key)
        try:
            agg[cur_line][1] += int(r[ci]); agg[cur_line][2] += int(r[ii])
        except ValueError:
            pass
tot = sum(v[1] for v in agg.values()); toti = sum(v[2] for v in agg.values())
print("samples", tot, "instructions", toti)
by = 2 if (len(sys.argv) > 3 and sys.argv[3] == "ins") else 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:topn]:
    print("%5.1f%% smp %5.1f%% ins  %s:%s  %s" % (100 * v[1] / max(tot, 1), 100 * v[2] / max(toti, 1), k[0], k[1], v[0].strip()[:110]))
