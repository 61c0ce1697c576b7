"""How much of a bench step is not GEMM kernel time?  Runs the bench's device step with per-launch profiling off and
on and prints wall / event times (dev aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv, _lib  # noqa: E402

P, K = 16384, 100
nu = np.linspace(bench.NU[0], bench.NU[1], bench.M_CH)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
src.sigma = np.full(bench.M_CH, bench.SIGMA)
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
plan = dv.DevicePlan.get(src.lambda2, src.l2_ref, par.phi)
dev = plan.device
data = dv.pack_planar(bench.make_slab(torch, src, P, 1234, dev), plan.m, plan.ld2m, dev)[0]
w = torch.as_tensor(src.w, dtype=torch.float32, device=dev)
s = torch.as_tensor(src.s, dtype=torch.float32, device=dev)
k = torch.full((P,), float(src.k), dtype=torch.float32, device=dev)
lam0 = 40.0 * src.theo_noise
lam = torch.full((P,), lam0, dtype=torch.float32, device=dev)
nz = torch.full((P,), lam0 / (2 * K), dtype=torch.float32, device=dev)
x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=dev)
lib = _lib.load()


def run(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        plan.fista(data, w, s, k, x, lam, nz, K, use_dirty_as_guess=True, want_delta=True)
    t_host = time.perf_counter() - t0      # host time to ENQUEUE everything
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * t_host / n


run(2)
for prof in (0, 1, 0):
    lib.csr_profile(prof)
    dev_ms, host_ms = run(3)
    print("profile=%d: device %.1f ms/step, host enqueue %.1f ms/step" % (prof, dev_ms, host_ms))
    if prof:
        ms, cnt = np.zeros(16), np.zeros(16, dtype=np.int64)
        import ctypes
        lib.csr_profile_collect(ms.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
        print("  kernel ms per step by kind:", np.round(ms / 3, 2), "sum %.1f" % (ms.sum() / 3))
