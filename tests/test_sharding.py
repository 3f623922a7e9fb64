"""CPU tests of the N>1 host path: frame sharding + ordered gather over torch.distributed (gloo, world_size 2).
The GPU stage is replaced by a deterministic stand-in (`compute=`), so this checks exactly the logic that
differs from the single-GPU path: which rank gets which frames, and that rank 0 writes results.txt / errors.txt
in the order a sequential run would (detecttrails.py:349-407)."""
import io
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from lfd_b200 import sharding


def test_shard_indices_partition():
    for n in (0, 1, 5, 16, 33, 100):
        for world in (1, 2, 3, 8):
            for block in (1, 4, 16):
                parts = [sharding.shard_indices(n, r, world, block) for r in range(world)]
                flat = sorted(i for p in parts for i in p)
                assert flat == list(range(n))
                for p in parts:
                    assert p == sorted(p)
                # consecutive frames of a block stay on one rank
                for r, p in enumerate(parts):
                    for i in p:
                        assert (i // block) % world == r


def test_merge_records_detects_holes_and_duplicates():
    assert sharding.merge_records(3, [([0, 2], ["a", "c"]), ([1], ["b"])]) == ["a", "b", "c"]
    with pytest.raises(ValueError):
        sharding.merge_records(3, [([0, 2], ["a", "c"])])
    with pytest.raises(ValueError):
        sharding.merge_records(2, [([0, 1], ["a", "b"]), ([1], ["b"])])


def _fake_compute(frames):
    out = []
    for (run, camcol, flt, field) in frames:
        if field % 7 == 0:
            out.append(("err", "%d %d %s %d\nboom\n\n" % (run, camcol, flt, field)))
        elif field % 3 == 0:
            out.append(("line", "%d %d %s %d 1 2 3 4\n" % (run, camcol, flt, field)))
        else:
            out.append(("none", ""))
    return out


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from lfd_b200.detecttrails import process_fields
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        frames = [(2888, 1, "r", f) for f in range(100, 137)]
        res, err = io.StringIO(), io.StringIO()
        process_fields(res, err, frames, {"debug": False}, {"debug": False}, {}, batch=4, compute=_fake_compute)
        with open(os.path.join(tmp, "out%d.txt" % rank), "w") as f:
            f.write(res.getvalue() + "---\n" + err.getvalue())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_ordered_gather(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    frames = [(2888, 1, "r", f) for f in range(100, 137)]
    recs = _fake_compute(frames)
    exp = "".join(t for k, t in recs if k == "line") + "---\n" + "".join(t for k, t in recs if k == "err")
    assert (tmp_path / "out0.txt").read_text() == exp          # rank 0 wrote everything, in sequential order
    assert (tmp_path / "out1.txt").read_text() == "---\n"      # other ranks write nothing


def test_resume_skips_finished_frames(tmp_path):
    """N3: an interrupted run restarted with the same progress file processes only what is left and appends nothing twice."""
    from lfd_b200.detecttrails import process_fields, read_progress
    frames = [(2888, 1, "r", f) for f in range(100, 160)]
    calls = []

    def compute(fr):
        calls.append(list(fr))
        return _fake_compute(fr)

    class Boom(Exception):
        pass

    def compute_crashing(fr):
        if any(f[3] >= 140 for f in fr):
            raise Boom()
        return compute(fr)

    prog = str(tmp_path / "progress.txt")
    res_p, err_p = tmp_path / "results.txt", tmp_path / "errors.txt"
    with open(res_p, "a") as res, open(err_p, "a") as err:
        with pytest.raises(Boom):
            process_fields(res, err, frames, {"debug": False}, {"debug": False}, {}, batch=1, compute=compute_crashing,
                           distributed=False, progress=prog)
    done = read_progress(prog)
    assert 0 < len(done) < len(frames)
    with open(res_p, "a") as res, open(err_p, "a") as err:
        process_fields(res, err, frames, {"debug": False}, {"debug": False}, {}, batch=1, compute=compute,
                       distributed=False, progress=prog)
    recs = _fake_compute(frames)
    assert res_p.read_text() == "".join(t for k, t in recs if k == "line")
    assert err_p.read_text() == "".join(t for k, t in recs if k == "err")
    seen = [f for c in calls for f in c]
    assert sorted(seen) == sorted(frames) and len(seen) == len(set(seen))      # every frame computed exactly once
    assert len(read_progress(prog)) == len(frames)


def test_spread_device_placement(monkeypatch):
    """More GPUs than ranks: ranks are spread over the device indices (host->device PCIe is the scarce link and
    neighbouring GPUs share a host bridge); as many ranks as GPUs, or counts that do not divide: identity."""
    from lfd_b200.sharding import spread_device
    monkeypatch.delenv("LFD_DEVICE_MAP", raising=False)
    assert [spread_device(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]
    assert [spread_device(r, 2, 8) for r in range(2)] == [0, 4]
    assert [spread_device(r, 8, 8) for r in range(8)] == list(range(8))
    assert [spread_device(r, 3, 8) for r in range(3)] == [0, 1, 2]
    assert spread_device(0, 1, 8) == 0 and spread_device(0, 1, 1) == 0
    monkeypatch.setenv("LFD_DEVICE_MAP", "packed")
    assert [spread_device(r, 4, 8) for r in range(4)] == [0, 1, 2, 3]


def test_h2d_gate_policy(monkeypatch):
    """Two copy slots per host bridge (devices 4k .. 4k+3) once more than two ranks share one; none otherwise."""
    from lfd_b200.sharding import h2d_gate_path
    monkeypatch.delenv("LFD_H2D_GATE", raising=False)
    monkeypatch.delenv("LFD_DEVICE_MAP", raising=False)
    paths = [h2d_gate_path(d, 8, 8) for d in range(8)]
    assert all(paths) and paths[0] == paths[2] != paths[1] == paths[3] and paths[4] == paths[6] != paths[5] == paths[7]
    assert len(set(paths)) == 4 and paths[0] != paths[4]
    assert [h2d_gate_path(d, 4, 8) for d in (0, 2, 4, 6)] == [None] * 4        # spread: two ranks per bridge
    assert h2d_gate_path(0, 1, 8) is None and h2d_gate_path(0, 2, 8) is None
    assert all(h2d_gate_path(d, 4, 4) for d in range(4))                         # four ranks on one bridge
    monkeypatch.setenv("LFD_H2D_GATE", "0")
    assert h2d_gate_path(0, 8, 8) is None


def test_apply_h2d_gate_degrades(monkeypatch):
    """A lock file that cannot be opened leaves the handles ungated (warning), it does not abort the run."""
    import warnings
    from lfd_b200.sharding import apply_h2d_gate, h2d_gate_path
    monkeypatch.delenv("LFD_H2D_GATE", raising=False)
    monkeypatch.delenv("LFD_DEVICE_MAP", raising=False)

    class Fake:
        def __init__(self, fail):
            self.fail, self.calls = fail, []

        def set_h2d_gate(self, path):
            self.calls.append(path)
            if path and self.fail:
                raise OSError("cannot open " + path)

    good = [Fake(False), Fake(False)]
    path = apply_h2d_gate(good, 1, 8, 8)
    assert path == h2d_gate_path(1, 8, 8) and all(h.calls == [path] for h in good)
    assert apply_h2d_gate(good, 0, 2, 8) is None and all(len(h.calls) == 1 for h in good)     # no gate needed: untouched
    bad = [Fake(False), Fake(True)]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert apply_h2d_gate(bad, 1, 8, 8) is None
    assert w and "gate disabled" in str(w[0].message)
    assert bad[0].calls[-1] is None and bad[1].calls[-1] is None
