"""small fixed workload for ncu: one fused FISTA call (C3 sampling, P lines of sight, K iterations)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nu = np.linspace(1.008e9, 2.031e9, 1000)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
plan = dv.DevicePlan.get(src.lambda2, 0.0, par.phi)
dev = plan.device
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the bench's own synthetic slab: sources at |phi| < 500 + noise)
src.sigma = np.full(1000, bench.SIGMA)
data = dv.pack_planar(bench.make_slab(torch, src, P, 1234, dev), plan.m, plan.ld2m, dev)[0]
w = torch.full((plan.m,), 1.0 / 0.0007 ** 2, device=dev)
s = torch.as_tensor(np.ascontiguousarray(src.s), dtype=torch.float32, device=dev)
k = torch.full((P,), float((w / s).sum()), device=dev)
x = torch.zeros((P, plan.ld2n), device=dev)
theo = 1.0 / np.sqrt(plan.m / 0.0007 ** 2)
lam = torch.full((P,), 40 * theo, device=dev)
nz = torch.full((P,), 40 * theo / 200, device=dev)
plan.fista(data, w, s, k, x, lam, nz, K, use_dirty_as_guess=True)
torch.cuda.synchronize()
print("ok", float(x.abs().max()))
