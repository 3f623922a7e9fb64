#!/usr/bin/env python
"""One frame list, all GPUs of the node: torchrun --nproc-per-node N examples/run_sharded.py <out_dir>

Builds a small synthetic SDSS tree, initialises torch.distributed (NCCL, one rank per GPU) and runs the ordinary
drop-in call; `process()` shards the frame list over the ranks and rank 0 writes results.txt / errors.txt in the
order a single-GPU run would (lfd_b200/sharding.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import lfd_b200
from lfd_b200 import synth

out = sys.argv[1] if len(sys.argv) > 1 else "."
from lfd_b200.sharding import spread_device

local = spread_device(os.environ.get("LOCAL_RANK", 0))     # the device DetectTrails picks too: ranks spread over the host bridges
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = os.path.join(out, "tree")
if dist.get_rank() == 0:
    tree = synth.write_sdss_tree(root, 2888, 1, range(100, 140), filters=("r",), startfield=100, endfield=140)
dist.barrier()
boss = os.path.join(root, "boss")
lfd_b200.setup(boss, os.path.join(boss, "photoObj"), os.path.join(boss, "photo", "redux"), out)
lfd_b200.DetectTrails(run=2888, camcol=1, filter="r", savepath=out, batch=8).process()
dist.barrier()
if dist.get_rank() == 0:
    print("results lines:", sum(1 for _ in open(os.path.join(out, "results.txt"))))
dist.destroy_process_group()
