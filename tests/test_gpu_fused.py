"""The fused, TMA-fed morphology + Sobel + NMS kernel (lfd_b200/csrc/k_mnms.cuh) is opt-in (LFD_FUSED=1, read when a
handle is created), so its parity is checked in a child process with the switch set: the stage-by-stage taps of both
passes (TAP = true instantiations: morph, eroded, NMS classes, Canny edges, box image, accumulators, lines) and a
production batch without taps (TAP = false, CUDA graph) against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from lfd_b200 import _lib, synth
from lfd_b200.removestars import star_rects
from oracle import ref_pipeline as rp
from oracle.verdicts import device_verdict, verdicts
kinds = [("trail", 1), ("sparse", 2), ("dense_trail", 3), ("satellite", 4), ("empty", 5), ("dense", 6), ("faint_trail", 7),
         ("trail_var", 8), ("trail_axis", 9), ("dense_heavy", 10), ("sparse", 11), ("trail", 12), ("satellite", 13),
         ("sparse", 14), ("trail_var", 15), ("sparse", 16), ("dense", 17), ("trail_axis", 18)]
frames, cats = zip(*[synth.make_case(k, s) for k, s in kinds])
pr = dict(rp.DEFAULT_REMOVESTARS)
rects = [star_rects(c, "r", f.shape, **pr) for f, c in zip(frames, cats)]
ref = verdicts(frames, cats, ["r"] * len(frames))
h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=len(frames))
h.set_params(dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM))
# production path: no tap flags, n >= 16 -> graph + batch parts; the fused kernel never writes the morph plane
h.submit(np.stack(frames), rects)
res = h.wait()
assert all(ms == 0.0 for _n, ms in h.timings()[2:]), "not the CUDA-graph path"
bad = [i for i in range(len(frames)) if device_verdict(res[i], frames[0].shape) != tuple(ref[i])]
assert not bad, bad
try:
    h.stage(0, 0, "morph")
    raise SystemExit("the morph plane should not exist without LFD_KEEP_TAPS when the fused kernel runs")
except _lib.LfdError:
    pass
# taps of a few frames, both passes, against the oracle's stage images
for i in (0, 2, 5, 9):
    taps = {}
    rp.process_frame(frames[i].copy(), cats[i], "r", taps=taps)
    h.submit(frames[i][None], [rects[i]], flags=_lib.KEEP_TAPS | _lib.SERIAL_PASSES)
    h.wait()
    for p, key in ((0, "bright"), (1, "dim")):
        t = taps[key]
        if not t:
            continue
        for stage, name in (("morph", "morph"), ("canny", "canny"), ("box", "box_img")):
            assert np.array_equal(h.stage(0, p, stage), t[name]), (i, key, stage)
        if p == 1:
            assert np.array_equal(h.stage(0, 1, "eroded"), t["eroded"]), (i, "eroded")
h.close()
print("FUSED-OK", sum(1 for r in ref if r[0] is True))
'''


def test_fused_kernel_parity_in_child_process():
    env = dict(os.environ, LFD_FUSED="1")
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert "FUSED-OK" in out.stdout
