#!/usr/bin/env python
"""bench.py - frames/s of lfd.detecttrails' full per-frame detect on synthetic SDSS-shaped frames.

  python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path on the host cores)

A step = one batch of B frames (2048x1489 float32, config-3 mix of sparse / dense / trail / satellite
fields with star catalogs) through star mask -> flip -> bright -> dim-if-needed -> result.
`value` is measured with the batch already resident in HBM; `e2e` is the same metric through the C-ABI
with HOST (pinned) frames, host->device and device->host copies inside the timed region, two handles
double-buffered so copies overlap compute.  Frames are independent: ranks shard frames, no collective.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lfd_b200 import synth  # noqa: E402

H, W = synth.FRAME_H, synth.FRAME_W
N = H * W
METRIC = "frames/sec (2048x1489 SDSS field, full detect)"
WORKLOAD = "config3-camcol-mix"


def env_int(k, d):
    return int(os.environ.get(k, d))


def make_pool(n, rank):
    """n distinct frames of the config-3 mix (seeded by run/camcol/filter/field), with their blot rects."""
    from lfd_b200.removestars import star_rects
    from oracle.ref_pipeline import DEFAULT_REMOVESTARS
    frames, cats, rects, kinds = [], [], [], []
    field = 11
    while len(frames) < n:
        flt = synth.FILTERS[len(frames) % 5]
        kind, seed = synth.case_for_frame(2888, 1 + rank % 6, flt, field + 1000 * (rank // 6))
        img, cat = synth.make_case(kind, seed)
        frames.append(img); cats.append(cat); kinds.append(kind)
        rects.append(star_rects(cat, flt, img.shape, **DEFAULT_REMOVESTARS))
        field += 1
    return frames, cats, rects, kinds


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p, self.f = gpu, None, None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f:
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); smax.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# CPU reference leg: the oracle's restatement of the reference's call sequence on cv2, all host cores
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_pipeline as rp
    img, cat, flt = args
    t = time.perf_counter()
    rp.process_frame(img.copy(), cat, flt)
    return time.perf_counter() - t


def cpu_frames_per_s(frames, cats, nframes, cores, repeat=1):
    """Wall-clock frames/s of the oracle pipeline over `nframes` frames on `cores` processes."""
    import multiprocessing as mp
    jobs = [(frames[i % len(frames)], cats[i % len(frames)], synth.FILTERS[i % 5]) for i in range(nframes)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, jobs[:cores])          # warm the workers (imports, page-in)
        best = None
        for _ in range(repeat):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return nframes / best, best


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = max(2 * cores, 8)
    frames, cats, _rects, _kinds = make_pool(min(per_step, 32), 0)
    import multiprocessing as mp
    jobs = [(frames[i % len(frames)], cats[i % len(frames)], synth.FILTERS[i % 5]) for i in range(per_step)]
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, jobs[:cores])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    import cv2
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [H, W], "frames_per_step": per_step, "cv2": cv2.__version__, "numpy": np.__version__},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames/step x %d steps of the %s pool, oracle/ref_pipeline.py (reference call sequence on cv2 %s), "
                                   "multiprocessing.Pool(%d), cv2.setNumThreads(1)" % (per_step, args.steps, WORKLOAD, cv2.__version__, cores)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lfd_b200 import _lib
    from oracle import ref_pipeline as rp

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    distributed = world > 1
    torch.cuda.set_device(local)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    frames, cats, rects, kinds = make_pool(B, rank)
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    hA = _lib.Handle(H, W, max_batch=B, device=local)
    hB = _lib.Handle(H, W, max_batch=B, device=local)
    for h in (hA, hB):
        h.set_params(pb, pd)
        for i, f in enumerate(frames):
            h.host_frames[i] = f                       # pinned host staging, filled once

    # ---- value: batch resident in HBM -------------------------------------------------------
    res = hA.upload(B, rects)                          # H2D once (+ one untimed run)
    n_detect = sum(r.detected for r in res)
    for _ in range(args.warmup):
        hA.run_resident(B); hA.wait()
    sampler = ClockSampler(local)
    stage_ms = None
    barrier()
    sampler.start()
    l0 = hA.kernel_launches()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        hA.run_resident(B); hA.wait()
        tm = hA.timings()
        dev_ms += sum(ms for _, ms in tm)
        stage_ms = tm if stage_ms is None else [(n, a + b) for (n, a), (_, b) in zip(stage_ms, tm)]
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    elapsed = time.perf_counter() - t0
    launches = hA.kernel_launches() - l0
    clocks = sampler.stop()
    elapsed = reduce_max(elapsed)
    dev_ms = reduce_max(dev_ms)
    counters = hA.counters()
    value = world * B * args.steps / elapsed

    if args.profile:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "profile_only": True,
                              "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B},
                              "stages": [{"stage": n, "ms_per_step": ms / args.steps} for n, ms in stage_ms]}))
        return 0

    # ---- e2e: host frames -> device -> results, double-buffered over two handles -------------
    def e2e_loop(steps):
        hs = (hA, hB)
        hs[0].submit(B, rects)
        for k in range(1, steps):
            hs[k & 1].submit(B, rects)
            hs[(k - 1) & 1].wait()
        hs[(steps - 1) & 1].wait()
    e2e_loop(max(args.warmup, 2))
    barrier()
    l1 = hA.kernel_launches() + hB.kernel_launches()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    e2e_elapsed = reduce_max(time.perf_counter() - t0)
    launches_e2e = hA.kernel_launches() + hB.kernel_launches() - l1
    e2e_value = world * B * args.steps / e2e_elapsed
    # H2D-only time (PCIe) reported separately
    t0 = time.perf_counter()
    for _ in range(3):
        hA._ck(hA._L.lfd_upload(hA.h, None, B, None, None, 0))
    torch.cuda.synchronize()
    import ctypes
    h2d_s = (time.perf_counter() - t0) / 3
    rect_bytes = int(sum(len(r) for r in rects) * 16 + (B + 1) * 4)
    h2d_bytes = B * N * 4 + rect_bytes
    d2h_bytes = B * (ctypes.sizeof(_lib.Result) + 96) + 128

    launches_total = int(reduce_sum(launches + launches_e2e))

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant HBM-bound kernel (k_prep) + per-stage report ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    stage_ms = [(n, ms / args.steps) for n, ms in stage_ms]
    per = dict(stage_ms)
    n_bright, n_dim = counters["frames_bright_run"], counters["frames_dim_run"]
    # algorithmic bytes: S1 = 4N read + 1N write per pass (both passes produced in one launch)
    prep_bytes = B * (4 * N + 2 * N)
    prep_ms = per["prep(blot+flip+clip+u8+hist)"]
    stage_report = []
    alg = {"lut+morph": 2 * N + 2 * N, "sobel+nms": 2 * N, "ccl_fg(hysteresis)": 9 * N // 2, "ccl_bg(holes)": 9 * N // 2,
           "rects+boxfill": N}
    for name, ms in stage_ms:
        entry = {"stage": name, "ms_per_step": round(ms, 4)}
        key = name.split(":")[-1]
        nfr = n_bright if name.startswith("bright") else n_dim if name.startswith("dim") else B
        if name.startswith("prep"):
            entry.update(bytes=prep_bytes, gbs=prep_bytes / (ms * 1e6) if ms > 0 else None)
        elif key in alg and ms > 0:
            by = nfr * alg[key]
            entry.update(bytes=by, gbs=by / (ms * 1e6))
        if entry.get("gbs"):
            entry["frac_of_hbm_peak"] = entry["gbs"] / peak
        stage_report.append(entry)
    hough_ms = per["bright:hough"] + per["dim:hough"]
    roofline = {"bound": "hbm", "kernel": "k_prep", "achieved": prep_bytes / (prep_ms * 1e6), "peak": peak, "unit": "GB/s",
                "frac": prep_bytes / (prep_ms * 1e6) / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": prep_bytes,
                "note": "k_prep reads the float frame once (4N) and writes both passes' uint8 planes (2N); "
                        "traffic from profiles/ ncu capture when present"}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            roofline["traffic"] = json.load(open(tr_path)).get("k_prep_bytes_per_launch_b%d" % B)
        except Exception:
            pass

    # ---- CPU baseline on this box's host cores (bounded sample) -------------------------------
    cores = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1:
        sample = min(max(2 * cores, 8), 128)
        cpu_value, cpu_s = cpu_frames_per_s(frames, cats, sample, cores)
        import cv2
        cpu_baseline = {"value": cpu_value, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": "%d frames of the same pool, oracle/ref_pipeline.py (reference call sequence on cv2 %s), "
                                  "multiprocessing.Pool(%d), cv2.setNumThreads(1), %.1f s" % (sample, cv2.__version__, cores, cpu_s)}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [H, W], "frames_per_step_per_gpu": B, "pool": "B distinct frames per rank "
                   "(sparse/dense/trail/satellite mix, seeded), inputs %.0f MB per step > L2 (126 MB), no L2 flush" % (B * N * 4 / 1e6),
                   "kinds": {k: kinds.count(k) for k in sorted(set(kinds))}, "detections_per_batch": int(n_detect),
                   "parallelism": "frame-sharded x%d, no collective" % world},
        "device_ms_per_step": dev_ms / args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_elapsed / args.steps, "h2d_only_ms_per_step": 1e3 * h2d_s,
                "h2d_gbs": h2d_bytes / h2d_s / 1e9, "how": "lfd_submit from pinned host staging + lfd_wait, two handles double-buffered"},
        "gpu_launches": launches_total,
        "roofline": roofline,
        "stages": stage_report,
        "hough": {"ms_per_step": hough_ms, "votes_per_step": counters["votes"],
                  "gvotes_per_s": counters["votes"] / (hough_ms * 1e6) if hough_ms > 0 else None,
                  "frames_hough": counters["frames_hough"]},
        "counters": counters,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile", action="store_true", help="device-resident leg only (target command for ncu)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
