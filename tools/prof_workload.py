"""fixed workload for ncu: one step of a bench.py workload (c3 / c4 / c5) at its bench slab, K iterations
    python tools/prof_workload.py c4 12 [slab]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from csromer_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = bench.Ctx()
ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.local = torch, None, 1, 0, 0
torch.cuda.set_device(0)
ctx.device = torch.device("cuda", 0)
ctx.lib = _lib.load()
slab = int(sys.argv[3]) if len(sys.argv) > 3 else bench.WORKLOADS[name]["slab"]
w = bench.Workload(ctx, name, slab, K)
torch.cuda.synchronize()
w.step(0, collective=False)
torch.cuda.synchronize()
print("ok", name, K, float(w.x.abs().max()))
