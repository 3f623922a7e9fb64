#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, ncu raw csv, plain_profile.json} into the tracked summaries.

usage: python profiles/summarize.py <round tag> <launches.csv> <raw.csv> <plain_profile.json> [batch]
  launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
  raw.csv      : ncu -i prof.ncu-rep --page raw --csv   ('-' = launch list only)
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, raw, plain = sys.argv[1:5]
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 16
P = os.path.join(ROOT, "profiles")

lines = [l for l in open(launches) if l.startswith('"')]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    tot[name] += float(row["Metric Value"])
    cnt[name] += 1
T = sum(tot.values())
pl = json.loads([l for l in open(plain) if l.startswith("{")][-1])
out_md = os.path.join(P, "%s_launch_shares.md" % tag)
with open(out_md, "w") as f:
    f.write("# %s: kernel shares of one step (batch %d frames)\n\n" % (tag, batch))
    f.write("Command: `python bench.py --profile --batch %d --steps 2 --warmup 3` (plain run first, exit 0, then under\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none`).  ncu times are cold-cache and serialised; "
            "compare shares.\n\n" % batch)
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for n, t in tot.most_common():
        f.write("| %s | %d | %.1f | %.1f | %.1f%% |\n" % (n, cnt[n], t / 1e3, t / cnt[n] / 1e3, 100 * t / T))
    f.write("\nCUDA-event stage times of the plain run (ms per step, %.0f frames/s):\n\n| stage | ms | share |\n|---|---:|---:|\n" % pl["value"])
    ts = sum(s["ms_per_step"] for s in pl["stages"])
    for s in pl["stages"]:
        f.write("| %s | %.3f | %.1f%% |\n" % (s["stage"], s["ms_per_step"], 100 * s["ms_per_step"] / ts))
print("wrote", out_md)
if raw == "-":              # launch list only (profiles/launch_list.sh)
    sys.exit(0)

rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(k) for k in keep if k in hdr]
out_csv = os.path.join(P, "%s_top_kernels_raw.csv" % tag)
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        rr = [r[i] for i in idx]
        rr[1] = rr[1].split("(")[0]
        w.writerow(rr)
print("wrote", out_csv)

mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = {}
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
for r in data:
    name = re.sub(r"<.*", "", r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")).strip()
    tr.setdefault(name, []).append(float(r[ir]) * mul[units[ir]] + float(r[iw]) * mul[units[iw]])
# batches of >= 16 frames run as two halves after k_prep, so every other launch of the profiled command covers
# half the batch
half = (batch + 1) // 2 if batch >= 16 else batch
traffic = {"batch": batch, "frames_per_launch": {"k_prep": batch, "others": half}, "source": os.path.basename(raw), "unit": "bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch"}
for k, v in tr.items():
    traffic["%s_bytes_per_launch_b%d" % (k, batch)] = sum(v) / len(v)
    traffic["%s_bytes_per_frame" % k] = sum(v) / len(v) / (batch if k == "k_prep" else half)
json.dump(traffic, open(os.path.join(P, "traffic_%s.json" % tag), "w"), indent=1)
print("wrote traffic_%s.json" % tag)
