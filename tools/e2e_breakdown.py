"""where does the end-to-end step spend its time? (dev aid; synchronises after every stage)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import csromer_b200 as cs  # noqa: E402

P, K = 16384, 100
dev = torch.device("cuda", 0)
nu = np.linspace(bench.NU[0], bench.NU[1], bench.M_CH)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
src.sigma = np.full(bench.M_CH, bench.SIGMA)
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
slab = bench.make_slab(torch, src, P, 1, dev).cpu().pin_memory()
lam0 = 40.0 * src.theo_noise
noise = lam0 / (2.0 * K)
out_host = torch.empty((2 * bench.N_PHI, P), dtype=torch.float32).pin_memory()


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("  %-34s %7.1f ms" % (label, 1e3 * (t1 - t0)))
    return t1


for rep in range(2):
    print("rep", rep)
    torch.cuda.synchronize()
    t = t_start = time.perf_counter()
    ds = cs.Dataset(lambda2=src.lambda2, data=slab, sigma=src.sigma, spectral_idx=1.0)
    dft = cs.NDFT1D(dataset=ds, parameter=par)
    t = tick("Dataset + NDFT1D objects", t)
    f_dirty = dft.backward(ds.data)
    t = tick("dft.backward (H2D + dirty)", t)
    guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
    guess.data = f_dirty
    guess.complex_data_to_real()
    t = tick("complex_data_to_real", t)
    chi2 = cs.Chi2(dft_obj=dft, F_dirty=f_dirty)
    l1 = cs.L1(reg=lam0)
    opt = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise, maxiter=K,
                   verbose=False)
    cost, X = opt.run()
    t = tick("FISTA.run (incl. model D2H)", t)
    out_host.copy_(X.data, non_blocking=True)
    t = tick("D2H solution", t)
    print("  total %.1f ms -> %.0f LOS/s" % (1e3 * (t - t_start), P / (t - t_start)))
