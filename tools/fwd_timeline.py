"""Role timeline of the K-block-skipping residual forward kernel (dft_gemm_kernel<EPI_RESID>) on a bench workload:
    CSR_TIMELINE=1 python tools/fwd_timeline.py c3 12
prints, per tile slot of the persistent CTAs, the median clock counts of: producer issue span, wait of the MMA thread for its first
operand, MMA issue span, epilogue wait for the accumulator (from its own start), epilogue body, and the tile period."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CSR_TIMELINE", "1")   # 2 = the three-product adjoint with the FISTA epilogue instead
import bench  # noqa: E402
from csromer_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = bench.Ctx()
ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.local = torch, None, 1, 0, 0
torch.cuda.set_device(0)
ctx.device = torch.device("cuda", 0)
ctx.lib = _lib.load()
w = bench.Workload(ctx, name, bench.WORKLOADS[name]["slab"], K)
torch.cuda.synchronize()
w.step(0, collective=False)
torch.cuda.synchronize()
n = 160 * 16 * 8
buf = np.zeros(n, dtype=np.int64)
got = ctx.lib.csr_debug_timeline(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), n)
t = buf[:got].reshape(-1, 16, 8)[:148]
print("CTAs with work:", int((t[:, 0, 6] > 0).sum()), " tiles per CTA (max):", int((t[:, :, 6] > 0).sum(axis=1).max()))
print("tile  n  kblocks | prod_span  mma_wait_first  mma_span | epi_wait  epi_body | period(epi end to end)  [median clocks; x0.51 ns at 1965 MHz]")
for k in range(10):
    v = t[:, k, :]
    ok = v[:, 6] > 0
    if not ok.any():
        break
    v = v[ok]
    per = (t[ok, k, 6] - t[ok, k - 1, 6]) if k else (v[:, 6] - v[:, 0])
    med = lambda a: int(np.median(a))
    print("%4d %3d %7.1f | %9d %15d %9d | %8d %9d | %8d" % (k, ok.sum(), np.mean(v[:, 7]), med(v[:, 1] - v[:, 0]), med(v[:, 2] - v[:, 0]),
          med(v[:, 3] - v[:, 2]), med(v[:, 5] - v[:, 4]), med(v[:, 6] - v[:, 5]), med(per)))
# how far ahead of the epilogue the other roles run: producer start of tile k+1 minus epilogue end of tile k
v0, v1 = t[:, 0, :], t[:, 1, :]
ok = (v1[:, 6] > 0)
if not ok.any():
    sys.exit(0)
print("producer start of tile 1 minus epilogue end of tile 0 (median):", int(np.median(v1[ok, 0] - v0[ok, 6])))
print("MMA last commit of tile 1 minus epilogue end of tile 0 (median):", int(np.median(v1[ok, 3] - v0[ok, 6])))
print("epilogue start of tile 1 minus epilogue end of tile 0 (median):", int(np.median(v1[ok, 4] - v0[ok, 6])))
