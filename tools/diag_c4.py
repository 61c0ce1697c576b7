"""diagnostic: how the screening cascade behaves on the C4 workload (per-pixel weights / flags); cumulative tensor-core
blocks by kind after K iterations, for variants of the slab"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import csromer_b200 as cs
from csromer_b200 import _device as dv

bench.set_workload("c4")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nu = np.linspace(bench.NU[0], bench.NU[1], bench.M_CH)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
src.sigma = np.full(bench.M_CH, bench.SIGMA)
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
plan = dv.DevicePlan.get(src.lambda2, src.l2_ref, par.phi)
dev = plan.device
s = torch.as_tensor(src.s, dtype=torch.float32, device=dev)
c, wt, pg = bench.make_slab_c4(torch, src, P, 1235, dev)
data0 = dv.pack_planar(c, plan.m, plan.ld2m, dev)[0]
plan.derotate(data0, pg)
x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=dev)
for name in ("c4", "no_flags", "const_sigma"):
    w = wt.clone()
    if name == "no_flags":
        w = torch.where(w == 0, torch.full_like(w, 1.0 / bench.SIGMA ** 2), w)
    if name == "const_sigma":
        w = torch.where(w == 0, w, torch.full_like(w, 1.0 / bench.SIGMA ** 2))
    k = (w / s[None, :]).sum(dim=1)
    lam = 40.0 / torch.sqrt(w.sum(dim=1))
    prev = np.zeros(4)
    for K in (1, 2, 5, 10, 20, 50, 100):
        nz = lam / 200.0
        out = plan.fista(data0.clone(), w, s, k, x, lam, nz, K, use_dirty_as_guess=True, want_mma_blocks=True)
        b = out["mma_blocks"].double().cpu().numpy()
        b = np.array([b[[0, 1, 2, 5]].sum(), b[8] + b[10], b[9], 0.0])  # three-product, fp16 one-product, fp8
        print(name, "K=%3d" % K, "blocks 3prod %.3g fp16 %.3g fp8 %.3g | last-range per-iter: 3prod %.3g fp16 %.3g fp8 %.3g" % (
            b[0], b[1], b[2], *(b[:3] - prev[:3])), "nnz", int((x != 0).sum()) // P)
        prev = b
