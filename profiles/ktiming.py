#!/usr/bin/env python
"""Per-launch device times of one resident step in real (warm, back-to-back) conditions, via LFD_KTIMING=1.
usage (GPU box): LFD_KTIMING=1 python profiles/ktiming.py [batch]"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LFD_KTIMING"] = "1"
import bench  # noqa: E402
from lfd_b200 import _lib  # noqa: E402
import lfd_b200  # noqa: E402
PB, PD, PR = lfd_b200.default_params()
PR = {k: v for k, v in PR.items() if k != "debug"}

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames, cats, rects, kinds = bench.make_pool(B, 0)
if os.environ.get("KIND"):           # e.g. KIND=sparse / KIND=dense: a batch of one kind of field only
    from lfd_b200 import synth
    from lfd_b200.removestars import star_rects
    frames, rects = [], []
    for i in range(B):
        img, cat = synth.make_case(os.environ["KIND"], 9000 + i % 8)
        frames.append(img); rects.append(star_rects(cat, "r", img.shape, **PR))
h = _lib.Handle(bench.H, bench.W, max_batch=B)
h.set_params(PB, PD)
for i, f in enumerate(frames):
    h.host_frames[i] = f
h.upload(B, rects)
src = open(os.path.join(ROOT, "lfd_b200", "csrc", "lfd_b200.cu")).read().split("\n")
acc = collections.OrderedDict()
R = 5
for _ in range(3):
    h.run_resident(B); h.wait()
for _ in range(R):
    h.run_resident(B); h.wait()
    seen = collections.Counter()
    for line, ms in h.ktimings():
        seen[line] += 1
        key = (line, seen[line])
        acc[key] = acc.get(key, 0.0) + ms
print("counters", h.counters())
tot = sum(acc.values()) / R
print("total %.3f ms / step of %d frames" % (tot, B))
for (line, k), ms in acc.items():
    m = None
    for back in range(1, 9):                 # the launch (or the macro that holds it) is at most a few lines above
        m = re.search(r"(k_\w+)", src[line - back])
        if m:
            break
    print("%-28s line %4d #%d  %8.1f us  %5.1f%%" % (m.group(1) if m else "?", line, k, 1e3 * ms / R, 100 * ms / R / tot))
