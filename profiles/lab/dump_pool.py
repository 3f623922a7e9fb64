#!/usr/bin/env python
"""Write the bench frame pool (float32 frames + the star mask as bit words, flipped orientation) to /tmp for prep_lab.
usage: python profiles/lab/dump_pool.py [batch]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames, cats, rects, kinds = bench.make_pool(B, 0)
if os.environ.get("KIND"):           # a batch of one kind of field only (sparse / dense / trail ...)
    import lfd_b200
    from lfd_b200 import synth
    from lfd_b200.removestars import star_rects
    PR = {k: v for k, v in lfd_b200.default_params()[2].items() if k != "debug"}
    frames, rects, kinds = [], [], []
    for i in range(B):
        img, cat = synth.make_case(os.environ["KIND"], 9000 + i % 8)
        frames.append(img); rects.append(star_rects(cat, "r", img.shape, **PR)); kinds.append(os.environ["KIND"])
H, W = frames[0].shape
WW = (W + 31) // 32
with open("/tmp/lab_in.bin", "wb") as fi, open("/tmp/lab_mask.bin", "wb") as fm:
    for f, rs in zip(frames, rects):
        fi.write(np.ascontiguousarray(f, np.float32).tobytes())
        m = np.zeros((H, WW * 32), bool)
        for r0, r1, c0, c1 in np.asarray(rs).reshape(-1, 4):
            m[max(r0, 0):min(r1, H), max(c0, 0):min(c1, W)] = True
        m = m[::-1]                                            # flipped orientation
        words = np.packbits(m.reshape(H, WW, 32), axis=2, bitorder="little").view(np.uint32).reshape(H, WW)
        fm.write(np.ascontiguousarray(words).tobytes())
print("wrote", B, "frames", kinds.count("dense"), "dense")
