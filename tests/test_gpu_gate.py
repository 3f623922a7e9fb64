"""lfd_set_h2d_gate: two processes that share one copy slot (an advisory lock file taken before a batch's H2D copy is
enqueued and released by a stream callback when the copy has completed) both finish, with the same results as ungated."""
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
gate = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else None
kinds = [("trail", 1), ("sparse", 2), ("satellite", 4), ("dense", 6)]
frames, cats = zip(*[synth.make_case(k, s) for k, s in kinds])
rects = [star_rects(c, "r", f.shape, **rp.DEFAULT_REMOVESTARS) for f, c in zip(frames, cats)]
hs = [_lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=4) for _ in range(2)]
for h in hs:
    h.set_params(dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM))
    h.set_h2d_gate(gate)
out = []
stack = np.stack(frames)
hs[0].submit(stack, rects)
for k in range(1, 12):                       # double-buffered like bench.py's e2e leg
    hs[k & 1].submit(stack, rects)
    out.append([bytes(r) for r in hs[(k - 1) & 1].wait()])
out.append([bytes(r) for r in hs[1].wait()])
assert all(o == out[0] for o in out)
for h in hs:
    h.set_h2d_gate(None)
    h.close()
import hashlib
print("GATE-OK", hashlib.sha1(b"".join(out[0])).hexdigest())
'''


def test_two_processes_share_a_copy_slot(tmp_path):
    gate = str(tmp_path / "slot.lock")
    code = CHILD % ROOT
    procs = [subprocess.Popen([sys.executable, "-c", code, gate], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
             for _ in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
        assert "GATE-OK" in so
    ref = subprocess.run([sys.executable, "-c", code, "-"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert ref.returncode == 0, ref.stderr[-2000:]
    digest = ref.stdout.split("GATE-OK")[1].strip()
    assert all(so.split("GATE-OK")[1].strip() == digest for so, _ in outs)
