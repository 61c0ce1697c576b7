"""Time the per-line-of-sight sampling kernels (K4, csrc/nudft.cu) with CUDA events: forward (dense and sparse
input), adjoint, and a K-iteration FISTA run, at the C1/C3 sampling (m=1000, n=7968) for P lines of sight that each
keep their own ~80 % subset of the channels.  Prints one JSON line (kept under profiles/).

usage: python tools/bench_k4.py [P=2048] [K=20]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m, keep = 1000, 0.8
rng = np.random.RandomState(0)
nu = np.linspace(1.008e9, 2.031e9, m)
full = cs.Dataset(nu=nu, sigma=np.ones(m), spectral_idx=1.0)
par = cs.Parameter()
par.calculate_cellsize(dataset=full, oversampling=8, verbose=False)
n = len(par.phi)
l2 = np.asarray(full.lambda2)
a = np.zeros((P, m))
cnt = np.zeros(P, dtype=np.int32)
for p in range(P):
    idx = np.sort(rng.choice(m, size=int(keep * m), replace=False))
    a[p, :idx.size] = l2[idx]
    cnt[p] = idx.size
st = dv.DeviceStack(a, cnt, par.phi)
dev = st.device
g = torch.Generator(device=dev).manual_seed(1)
x_dense = torch.randn((P, st.ld2n), generator=g, device=dev)
x_sparse = x_dense * (torch.rand((P, st.ld2n), generator=g, device=dev) < 0.002)
b = torch.randn((P, st.ld2m), generator=g, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


pairs = float(cnt.sum()) * n                     # (channel, phi cell) pairs per transform
res = {}
for name, fn in (("forward_dense", lambda: st.forward(x_dense)), ("forward_sparse_0.2pct", lambda: st.forward(x_sparse)),
                 ("adjoint", lambda: st.adjoint(b))):
    ms = timed(fn)
    res[name] = {"ms": ms, "algorithmic_tflops": 8.0 * pairs / ms / 1e9, "gpairs_per_s": pairs / ms / 1e6}

# FISTA on synthetic thin sources
w = torch.zeros((P, m), device=dev)
for p in range(P):
    w[p, :cnt[p]] = 1.0 / 0.0007 ** 2
s = torch.ones((P, m), device=dev)
k = (w / s).sum(dim=1)
at = torch.as_tensor(a, device=dev)
phi0 = torch.rand(P, generator=g, device=dev, dtype=torch.float64) * 1000 - 500
clean = 0.0035 * torch.exp(2j * phi0[:, None] * at)
data = torch.zeros((P, st.ld2m), device=dev)
data[:, :m] = (clean.real + 0.0007 * torch.randn((P, m), generator=g, device=dev, dtype=torch.float64)).float() * (w > 0)
data[:, m:2 * m] = (clean.imag + 0.0007 * torch.randn((P, m), generator=g, device=dev, dtype=torch.float64)).float() * (w > 0)
theo = 1.0 / torch.sqrt(w.sum(dim=1))
lam = (40 * theo).float()
nz = lam / (2 * K)
xs = torch.zeros((P, st.ld2n), device=dev)
ms = timed(lambda: st.fista(data, w, s, k, xs, lam, nz, K, use_dirty_as_guess=True), reps=2)
res["fista"] = {"iterations": K, "ms": ms, "ms_per_iteration": ms / K, "los_per_s_at_K100": P / (ms / K * 100) * 1e3,
                "finite": bool(torch.isfinite(xs).all()), "nonzero_fraction": float((xs != 0).float().mean())}
# fp32 pipe roofline: 148 SMs x 128 FMA lanes x clock; one pair costs >= 4 FFMA (complex multiply-add)
print(json.dumps({"workload": "K4 per-line-of-sight samplings: %d LOS, %d of %d channels each, n_phi=%d" % (P, int(keep * m), m, n),
                  "kernels": res,
                  "note": "FP32-pipe bound (4 FFMA per pair minimum; 148 SM x 128 lanes x ~1.9 GHz = 36 T pairs-FFMA/s -> "
                          "9 T pairs/s ceiling); HBM traffic is 8(m+n) B per line of sight per transform"}))
