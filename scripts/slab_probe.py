"""torchrun script: time the slab x sweeps with peer reads / writes selectively made local."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glia_b200.rd import RDHandle
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
def ag(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = RDHandle(n, "f32", device=local, dt_ctx=0.1, rank=rank, nranks=world, all_gather=ag)
F = 4.0 * n ** 3 / world
for what, name in ((0, "pc x sweep"), (1, "D-apply x sweep")):
    for mask, mn in ((0, "pull+push"), (1, "push only (reads local)"), (2, "pull only (writes local)"), (3, "all local")):
        ms = h.probe_xsweep(what, mask, 20)
        if rank == 0:
            nv = F * (world - 1) / world
            print(f"{name:16s} {mn:26s} {ms*1e3:8.1f} us   remote bytes/dir {nv/1e6:7.1f} MB -> {nv/ms/1e6:7.1f} GB/s if that were the bound", flush=True)
h.close()
dist.destroy_process_group()
