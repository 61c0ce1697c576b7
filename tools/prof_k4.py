"""fixed K4 workload for ncu: python tools/prof_k4.py (1024 lines of sight, 6 iterations)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from csromer_b200 import _lib
class A: pass
args = A(); args.steps = 1; args.sub_steps = 1; args.iters = 6; args.no_parity = True
ctx = bench.Ctx(); ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.local = torch, None, 1, 0, 0
ctx.barrier = torch.cuda.synchronize
torch.cuda.set_device(0); ctx.device = torch.device("cuda", 0); ctx.lib = _lib.load()
print(str(bench.run_k4(ctx, args, {"hbm": 6539.0}, P=1024))[:200])
