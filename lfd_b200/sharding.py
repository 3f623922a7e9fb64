"""Frame sharding across GPUs: frames are independent, so the ordered frame list is dealt to
the ranks in blocks and only the per-frame result records (<= ~100 bytes each) travel back - a host-side ordered
gather on rank 0, no collective on the data path and no NVLink traffic.

This replaces the reference's only scale-out story, hand-submitted PBS jobs whose results.txt files are
concatenated afterwards (/root/reference/lfd/createjobs/createjobs.py:173-218, lfd/createjobs/generic), and keeps
what those jobs cannot: results.txt / errors.txt in exactly the order a sequential run writes them
(/root/reference/lfd/detecttrails/detecttrails.py:349-407)."""


def is_distributed():
    try:
        import torch.distributed as dist
    except ImportError:      # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    import torch.distributed as dist
    return dist.get_rank() if is_distributed() else 0


def world_size():
    import torch.distributed as dist
    return dist.get_world_size() if is_distributed() else 1


def spread_device(local_rank, local_world=None, ngpu=None):
    """CUDA device for a local rank when the node has more GPUs than ranks: ranks are spread evenly over the device
    indices (rank r of N on G GPUs -> device r * (G // N)) instead of packed on devices 0..N-1.  The pipeline's scarce
    link is host->device PCIe, and GPUs with neighbouring indices share a host bridge on the 8 x B200 boxes this was
    measured on (profiles/h2d_lab_*.txt: devices 0-3 share ~120 GB/s, 0-7 ~235 GB/s, any two ~55 GB/s each), so four
    ranks on devices 0, 2, 4, 6 get twice the copy bandwidth of four ranks on 0-3.  With as many ranks as GPUs (or when
    the counts do not divide) this is the identity.  LFD_DEVICE_MAP=packed keeps device = local rank."""
    import os
    local_rank = int(local_rank)
    if os.environ.get("LFD_DEVICE_MAP", "spread") == "packed":
        return local_rank
    try:
        if local_world is None:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", 1)))
        if ngpu is None:
            import torch
            ngpu = torch.cuda.device_count()
    except Exception:   # noqa: BLE001
        return local_rank
    if local_world < 1 or ngpu <= local_world or ngpu % local_world != 0:
        return local_rank
    return local_rank * (ngpu // local_world)


BRIDGE_GPUS = 4      # GPUs per PCIe host bridge on the 8 x B200 boxes measured (profiles/h2d_lab_subsets.txt)
BRIDGE_SLOTS = 2     # concurrent batch copies a bridge serves at full speed (two GPUs: 55.5 GB/s each; four: 21-37 GB/s)


def h2d_gate_path(device, local_world=None, ngpu=None):
    """Lock file of the host->device copy slot this rank shares with its bridge neighbours, or None when no gate is needed.

    When more than BRIDGE_SLOTS ranks of this node use GPUs of one host bridge (devices 4k .. 4k+3), their concurrent
    781 MB batch copies are served unfairly and the slowest rank sets the pace of an equal-work job; the ranks of a bridge
    are then dealt round-robin to BRIDGE_SLOTS slots and `Handle.set_h2d_gate` makes the ranks of a slot take turns.
    LFD_H2D_GATE=0 disables this, LFD_H2D_GATE=1 forces a gate even for one or two ranks per bridge."""
    import os
    import tempfile
    mode = os.environ.get("LFD_H2D_GATE", "auto")
    if mode == "0":
        return None
    try:
        if local_world is None:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", 1)))
        if ngpu is None:
            import torch
            ngpu = torch.cuda.device_count()
    except Exception:   # noqa: BLE001
        return None
    devs = [spread_device(r, local_world, ngpu) for r in range(local_world)]
    bridge = int(device) // BRIDGE_GPUS
    on_bridge = sorted(d for d in devs if d // BRIDGE_GPUS == bridge)
    if int(device) not in on_bridge or (len(on_bridge) <= BRIDGE_SLOTS and mode != "1"):
        return None
    slot = on_bridge.index(int(device)) % BRIDGE_SLOTS
    return os.path.join(tempfile.gettempdir(), "lfd_b200_h2d_gate_%d_%d_%d" % (os.getuid(), bridge, slot))


def apply_h2d_gate(handles, device, local_world=None, ngpu=None):
    """Give `handles` the copy slot of `device` (h2d_gate_path); returns the lock path in use or None.  The gate is an
    optimisation: a lock file that cannot be opened (read-only or foreign /tmp) leaves the handles ungated with a warning."""
    gate = h2d_gate_path(device, local_world, ngpu)
    if not gate:
        return None
    try:
        for h in handles:
            h.set_h2d_gate(gate)
    except Exception as e:   # noqa: BLE001
        import warnings
        warnings.warn("lfd_b200: host->device copy gate disabled (%s)" % e)
        for h in handles:
            try:
                h.set_h2d_gate(None)
            except Exception:   # noqa: BLE001
                pass
        return None
    return gate


def shard_indices(n, rank, world, block):
    """Indices of the frames rank `rank` processes: blocks of `block` consecutive frames, round-robin over ranks
    (full GPU batches, and neighbouring fields - which share catalog/FITS directories - stay together)."""
    block = max(int(block), 1)
    out = []
    for b0 in range(rank * block, n, world * block):
        out.extend(range(b0, min(b0 + block, n)))
    return out


def merge_records(n, parts):
    """parts: iterable of (indices, records) per rank -> the n records in frame order."""
    merged = [None] * n
    for idxs, recs in parts:
        if len(idxs) != len(recs):
            raise ValueError("a rank returned %d records for %d frames" % (len(recs), len(idxs)))
        for i, r in zip(idxs, recs):
            if merged[i] is not None:
                raise ValueError("frame %d was processed twice" % i)
            merged[i] = r
    missing = [i for i, r in enumerate(merged) if r is None]
    if missing:
        raise ValueError("frames %s were not processed by any rank" % missing[:8])
    return merged


def run_sharded(frames, compute, block=16, group=None):
    """Each rank runs compute(its frames) and rank 0 returns the merged, ordered record list (other ranks: None)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    idxs = shard_indices(len(frames), rank, world, block)
    recs = compute([frames[i] for i in idxs])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((idxs, recs), gathered, dst=0, group=group)
    if rank != 0:
        return None
    return merge_records(len(frames), gathered)
