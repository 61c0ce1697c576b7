"""small fixed wavelet-basis workload for ncu: one fused FISTA call, SWT coif2; usage: prof_target_wav.py [P] [K] [nch]
(nch = 1000: C3 sampling, 4000: C5 sampling; thin sources at |phi| < 500 rad/m2 plus noise, as tools/bench_wavelet.py)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nch = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
band = (0.9e9, 1.67e9) if nch != 1000 else (1.008e9, 2.031e9)
nu = np.linspace(band[0], band[1], nch)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
plan = dv.DevicePlan.get(src.lambda2, 0.0, par.phi)
dev = plan.device
eng = cs.UndecimatedWavelet(wavelet_name="coif2").prepare(2 * plan.n)
g = torch.Generator(device=dev).manual_seed(1)
m = plan.m
l2 = torch.as_tensor(np.ascontiguousarray(src.lambda2), dtype=torch.float64, device=dev)[None, :]
phi = (torch.rand(P, generator=g, device=dev, dtype=torch.float64) - 0.5)[:, None] * 1000.0
amp = 0.0035 * (0.5 + torch.rand(P, generator=g, device=dev, dtype=torch.float64))[:, None]
data = torch.zeros((P, plan.ld2m), dtype=torch.float32, device=dev)
data[:, :m] = (amp * torch.cos(2 * phi * l2)).float() + 0.0007 * torch.randn((P, m), generator=g, device=dev)
data[:, m:2 * m] = (amp * torch.sin(2 * phi * l2)).float() + 0.0007 * torch.randn((P, m), generator=g, device=dev)
w = torch.full((plan.m,), 1.0 / 0.0007 ** 2, device=dev)
s = torch.as_tensor(np.ascontiguousarray(src.s), dtype=torch.float32, device=dev)
k = torch.full((P,), float((w / s).sum()), device=dev)
x = torch.zeros((P, eng.ldc), device=dev)
theo = 1.0 / np.sqrt(plan.m / 0.0007 ** 2)
lam = torch.full((P,), 40 * theo, device=dev)
nz = torch.full((P,), 40 * theo / 200, device=dev)
plan.fista(data, w, s, k, x, lam, nz, K, step=0.4, use_dirty_as_guess=True, wavelet=eng)
torch.cuda.synchronize()
print("ok", float(x.abs().max()))
