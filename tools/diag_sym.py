"""diagnostic: executed tensor-core blocks by kind with the mirror-symmetric screening on / off (C3 slab)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import csromer_b200 as cs
from csromer_b200 import _device as dv, _lib
P, K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 100
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
x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=dev)
for sym in (1, 0):
    _lib.set_option("sym", sym)
    prev = np.zeros(4)
    for kk in (2, 5, 10, 20, 50, 100):
        nz = torch.full((P,), lam0 / (2 * K), dtype=torch.float32, device=dev)
        out = plan.fista(data, w, s, k, x, lam, nz, kk, use_dirty_as_guess=True, want_mma_blocks=True)
        b = out["mma_blocks"].double().cpu().numpy()
        b = np.array([b[[0, 1, 2, 5]].sum(), b[8] + b[10], b[9], 0.0])  # three-product, fp16 one-product, fp8
        print("sym=%d K=%3d per-iteration in the last range: 3prod %.3g fp16 %.3g fp8 %.3g" % (sym, kk, *((b[:3] - prev[:3]) / (kk - (0 if prev[0] == 0 else kprev)))))
        prev, kprev = b, kk
