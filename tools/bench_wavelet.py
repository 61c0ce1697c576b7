"""Measurement of the WAVELET-basis configurations (BASELINE configs[4] sampling: 4000 channels, MeerKAT band,
n = 31968, undecimated coif2, L = 6; or the C3 sampling with --nch 1000), device-resident, per-kernel timing with
CUDA events (csr_profile).  Not the driver's bench line (bench.py measures configs[2]); results go to profiles/.

    python tools/bench_wavelet.py [--nch 4000] [--los 4096] [--iters 20] [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_wavelet.py --iters 100 --reps 2        # pixel-sharded: every rank its own slab, max over ranks
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv  # noqa: E402
from csromer_b200 import _lib  # noqa: E402
from csromer_b200.cube import estimate_lipschitz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nch", type=int, default=4000)
ap.add_argument("--los", type=int, default=4096)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

band = (0.9e9, 1.67e9) if args.nch != 1000 else (1.008e9, 2.031e9)
nu = np.linspace(band[0], band[1], args.nch)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
src.simulate()
src.sigma = np.full(args.nch, 0.0007)
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
dft = cs.NDFT1D(dataset=src, parameter=par)
plan = dft.device_plan()
dev = plan.device
m, n, P, K = plan.m, plan.n, args.los, args.iters
step = 1.0 / estimate_lipschitz(dft)
eng = cs.UndecimatedWavelet(wavelet_name="coif2").prepare(2 * n)
g = torch.Generator(device=dev).manual_seed(1236 + 17 * rank)
l2 = torch.as_tensor(np.ascontiguousarray(src.lambda2), dtype=torch.float64, device=dev)[None, :]
phi = (torch.rand(P, generator=g, device=dev, dtype=torch.float64) - 0.5)[:, None] * 1000.0
amp = 0.0035 * (0.5 + torch.rand(P, generator=g, device=dev, dtype=torch.float64))[:, None]
data = torch.zeros((P, plan.ld2m), dtype=torch.float32, device=dev)
data[:, :m] = (amp * torch.cos(2 * phi * l2)).float() + 0.0007 * torch.randn((P, m), generator=g, device=dev)
data[:, m:2 * m] = (amp * torch.sin(2 * phi * l2)).float() + 0.0007 * torch.randn((P, m), generator=g, device=dev)
w = torch.as_tensor(src.w, dtype=torch.float32, device=dev)
s = torch.as_tensor(np.ascontiguousarray(src.s), dtype=torch.float32, device=dev)
k = torch.full((P,), float(src.k), device=dev)
x = torch.zeros((P, eng.ldc), device=dev)
lam = torch.full((P,), 40 * src.theo_noise, device=dev)
nz = torch.full((P,), 40 * src.theo_noise / (2 * K), device=dev)
lib = _lib.load()


def run():
    return plan.fista(data, w, s, k, x, lam, nz, K, step=step, use_dirty_as_guess=True, wavelet=eng)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


run()
barrier()
lib.csr_profile(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
barrier()
lib.csr_profile(0)
ms = np.zeros(16)
cnt = np.zeros(16, dtype=np.int64)
_lib.check(lib.csr_profile_collect(ms.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                   cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))), "collect")
total_ms = e0.elapsed_time(e1) / args.reps
if world > 1:  # the slowest rank defines the step
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
C, N = eng.ncoef, eng.N
flops = 8.0 * m * n * P
names = ["plain_forward", "plain_adjoint", "resid_forward", "resid_adjoint", "fista_forward", "fista_adjoint",
         "swt_synth_fused", "swt_analysis_update_fused", "screen_adjoint_fp16", "screen_adjoint_fp8",
         "plain_adjoint_one_product", "swt_analysis_update_redo"]
screened = cnt[10] > 0  # wavelet-domain screening: the adjoint launches process tile lists, not the whole matrix
kern = {}
for i, name in enumerate(names):
    if cnt[i] == 0:
        continue
    avg = ms[i] / cnt[i]
    d = {"launches": int(cnt[i]), "avg_ms": float(avg), "share_of_run": float(ms[i] / (total_ms * args.reps))}
    if i >= 8 or (screened and i == 1):
        pass
    elif i < 6:
        d["algorithmic_tflops"] = flops / avg / 1e9
        d["frac_of_sustained_bf16_peak"] = d["algorithmic_tflops"] / peaks["bf16_tflops_sustained"]
        d["executed_frac"] = 3 * d["frac_of_sustained_bf16_peak"]
    else:
        nbytes = (C + N) * 4.0 * P if i == 6 else (N + 4 * C) * 4.0 * P
        d["algorithmic_bytes"] = nbytes
        d["achieved_gbs"] = nbytes / avg / 1e6
        d["frac_of_hbm_peak"] = d["achieved_gbs"] / peaks["hbm_gbs"]
    kern[name] = d
line = {"workload": "%d ch (%.3f-%.3f GHz), n_phi=%d, SWT coif2 L=%d (ncoef=%d), %d LOS, K=%d FISTA iterations, "
                    "step=1/lambda_max=%.3f, lambda0=40 theo_noise" % (m, band[0] / 1e9, band[1] / 1e9, n, eng.levels, C, P, K, step),
        "n_gpus": world, "los_per_gpu": P, "scaling": "weak (pixel-sharded, no data-path collective)",
        "ms_per_run": total_ms, "ms_per_iteration": total_ms / K, "los_per_s_at_K": world * P / total_ms * 1e3,
        "los_per_s_at_K100_extrapolated": world * P / (total_ms / K * 100) * 1e3,
        "full_c5_cube_seconds_at_this_rate": 2048 * 2048 / (world * P / (total_ms / K * 100) * 1e3),
        "finite": bool(torch.isfinite(x).all()), "kernels_rank0": kern}
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
