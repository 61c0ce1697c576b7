#!/usr/bin/env python
"""bench.py -- lines of sight per second for fixed-iteration FISTA (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference's CPU algorithm on the host cores)

Workload (config.workload): BASELINE.json configs[2] -- a 512 x 512-pixel cube x 1000 channels (JVLA band
1.008-2.031 GHz, n = 7968 Faraday-depth cells at oversampling 8), shared lambda^2 sampling, delta-basis L1 FISTA
with K_fista = 100 iterations, lambda0 = 40 theo_noise, noise = lambda0 / 200.  One STEP =
one pass of the whole hot path (dirty spectrum + 100 iterations + final evaluate = 2*100+2 transforms per line of
sight) over one slab of `--slab` lines of sight (default 16384 = 32 image rows); the cube is 16 such slabs.  With N
GPUs every rank runs its own slab per step (pixels shard with no data-path communication; weak scaling); the only
collective is the all-reduce of the convergence norm after each step.

value  : whole-job LOS/s with the slab resident in HBM (x, z state = 2 GB per slab >> 126 MB L2, so no L2 flush is
         needed between steps; steps alternate between two different slabs).
e2e    : the same metric through the public API (Dataset / NDFT1D / Chi2 / L1 / FISTA.run) with HOST buffers: the
         slab's data comes from pinned host memory every step and the solution cube and cost are read back.
roofline: the transform GEMM kernel (dft_gemm_kernel, both directions), algorithmic 8*m*n flops per line of sight
         per launch over CUDA-event launch durations recorded inside the timed region (csr_profile), against the
         measured dense bf16 tensor peak.  The kernel executes 3 fp16 products per algorithmic product (hi/lo
         split for fp32-class accuracy), so frac <= 1/3 by construction; `executed_tflops` is 3x achieved.
cpu_baseline: the oracle's NumPy port of the reference path (twiddles rebuilt with np.exp in every transform, as
         ndft.py:19-34 does), one line of sight per host core like the README's joblib recipe, bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_CH, N_PHI = 1000, 7968
NU = (1.008e9, 2.031e9)
SIGMA = 0.2 * 0.0035
WORKLOAD = "c3"
# --workload c4 (not the driver's line; results under profiles/): BASELINE.json configs[3] at slab scale -- 2000
# channels (n = 15968), per-pixel weights sigma0 U(0.5, 2) per channel, 10-20 % of each pixel's channels flagged
# (w = 0), per-pixel Galactic RM ~ N(0, 30) removed on the device inside the step (Dataset.subtract_galacticrm).
WORKLOADS = {"c3": dict(m=1000, n=7968, slab=16384), "c4": dict(m=2000, n=15968, slab=8192)}


def set_workload(name):
    global M_CH, N_PHI, WORKLOAD
    WORKLOAD = name
    M_CH, N_PHI = WORKLOADS[name]["m"], WORKLOADS[name]["n"]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------------------------
def _cpu_one_los(args):
    seed, k_sample = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        limiter = None
    from oracle import restatement as ob
    rng = np.random.RandomState(seed)
    nu = np.linspace(NU[0], NU[1], M_CH)
    l2 = ob.lambda2_from_nu(nu)
    sc = ob.dataset_scalars(l2, np.full(M_CH, SIGMA), 1.0)
    grid = ob.phi_grid(l2, sc["w"], 8.0)
    data = ob.thin_source(l2, 0.0035 * rng.uniform(0.5, 1.5), rng.uniform(-500, 500), 1.0)
    data = data + rng.normal(0, SIGMA, M_CH) + 1j * rng.normal(0, SIGMA, M_CH)
    op = ob.ExactDFT(l2, grid["phi"], sc["w"], sc["s"], sc["k"], rebuild_twiddles=True)
    t0 = time.perf_counter()
    x0 = ob.complex_to_real(op.backward(data))
    lam0 = 40.0 * sc["theo_noise"]
    ob.fista(op, data, x0, lam0, lam0 / 200.0, k_sample)
    dt = time.perf_counter() - t0
    del limiter
    return dt


def cpu_sample(k_sample, k_full, cores):
    """one line of sight per core, k_sample iterations each; returns (LOS/s extrapolated to k_full, wall, text)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        per_los = pool.map(_cpu_one_los, [(1000 + i, k_sample) for i in range(cores)])
    wall = time.perf_counter() - t0
    transforms_sample, transforms_full = 2 * k_sample + 2, 2 * k_full + 2
    busy = max(per_los)
    los_per_s = cores / (busy * transforms_full / transforms_sample)
    text = ("%d LOS (one per core, m=%d n=%d), %d of %d FISTA iterations each = %d of %d transforms, slowest worker "
            "%.2f s, extrapolated linearly in the transform count" %
            (cores, M_CH, N_PHI, k_sample, k_full, transforms_sample, transforms_full, busy))
    return los_per_s, wall, text


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_sample(1, args.iters, cores)
    t0 = time.perf_counter()
    vals = []
    text = ""
    for _ in range(args.steps):
        v, _, text = cpu_sample(args.cpu_iters, args.iters, cores)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "lines-of-sight/sec (fixed-iter FISTA)", "value": value, "unit": "LOS/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "LOS/s", "cores": cores, "kind": "port", "sample": text},
        "e2e": {"value": value, "unit": "LOS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ncu_traffic(P):
    """DRAM bytes per launch of the dominant kernel (the screening adjoint GEMM) from the committed `ncu --set full`
    capture (profiles/r1_traffic.json, taken at 16384 lines of sight per launch); None if the slab size differs or the
    file is missing."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if P != 16384 or WORKLOAD != "c3" or not os.path.exists(path):
        return None
    return json.load(open(path)).get("dominant_kernel_bytes")


def workload_config(args):
    text = ("C3: 512x512 px x 1000 ch cube (JVLA 1.008-2.031 GHz), shared lambda^2, n_phi=7968, delta-basis L1 FISTA, "
            "exact DFT operator")
    if WORKLOAD == "c4":
        text = ("C4: 1024x1024 px x 2000 ch cube (JVLA 1.008-2.031 GHz), n_phi=15968, per-pixel weights and 10-20 % "
                "flagged channels per pixel, per-pixel Galactic RM removed on the device inside the step, delta-basis L1 "
                "FISTA, exact DFT operator")
    return {"workload": text,
            "fista_iterations": args.iters, "lambda0": "40*theo_noise", "noise": "lambda0/200",
            "los_per_step_per_gpu": args.slab, "transforms_per_los": 2 * args.iters + 2,
            "l2_policy": "inputs larger than L2 (state 2 GB per slab; steps alternate between two slabs)",
            "e2e_pipeline": "public API per slab; the D2H of step i's solution overlaps step i+1's compute (second stream, "
                            "two pinned buffers); all copies inside the timed region",
            "parallelism": "pixel-sharded, no data-path collective"}


# ---------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------
# synthetic slab: generated on the device, outside every timed region
# ---------------------------------------------------------------------------------------------------------
def make_slab(torch, src, P, seed, device):
    """thin source per pixel (phi ~ U(-500, 500), amplitude 0.0035 U(0.5, 1.5)), 30 % of pixels + a thick slab
    (phi_fg ~ U(50, 200) centred on 200), alpha = 1, Gaussian noise sigma = 0.2 * 0.0035.  Returns complex64 (m, P)."""
    g = torch.Generator(device=device).manual_seed(seed)
    l2 = torch.as_tensor(np.ascontiguousarray(src.lambda2), dtype=torch.float64, device=device)[:, None]
    s = torch.as_tensor(np.ascontiguousarray(src.s), dtype=torch.float64, device=device)[:, None]
    phi = (torch.rand(P, generator=g, device=device, dtype=torch.float64) - 0.5) * 1000.0
    amp = 0.0035 * (0.5 + torch.rand(P, generator=g, device=device, dtype=torch.float64))
    data = amp * s * torch.exp(2j * phi * l2)
    thick = torch.rand(P, generator=g, device=device) < 0.3
    fg = 50.0 + 150.0 * torch.rand(P, generator=g, device=device, dtype=torch.float64)
    data = data + thick * (0.0035 * s * torch.exp(2j * l2 * 200.0) * torch.sinc(fg * l2 / np.pi))
    if WORKLOAD == "c4":
        return data, g
    noise = torch.randn((2, src.m, P), generator=g, device=device, dtype=torch.float64) * SIGMA
    return (data + torch.complex(noise[0], noise[1])).to(torch.complex64)


def make_slab_c4(torch, src, P, seed, device):
    """C4 on top of the C3 sources: per-pixel sigma and flags, Galactic RM screen.  Returns (complex64 (m, P) data,
    fp32 [P, m] weights, fp32 [P] phi_gal)."""
    clean, g = make_slab(torch, src, P, seed, device)
    m = src.m
    sigma = SIGMA * (0.5 + 1.5 * torch.rand((m, P), generator=g, device=device, dtype=torch.float64))
    frac = 0.1 + 0.1 * torch.rand((1, P), generator=g, device=device, dtype=torch.float64)
    flagged = torch.rand((m, P), generator=g, device=device, dtype=torch.float64) < frac
    noise = torch.randn((2, m, P), generator=g, device=device, dtype=torch.float64) * sigma
    phi_gal = 30.0 * torch.randn(P, generator=g, device=device, dtype=torch.float64)
    l2 = torch.as_tensor(np.ascontiguousarray(src.lambda2), dtype=torch.float64, device=device)[:, None]
    data = (clean + torch.complex(noise[0], noise[1])) * torch.exp(2j * phi_gal * l2)
    w = torch.where(flagged, torch.zeros_like(sigma), 1.0 / sigma ** 2)
    return data.to(torch.complex64), w.t().contiguous().float(), phi_gal.float()


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    ge.build()
    import csromer_b200 as cs
    from csromer_b200 import _device as dv
    from csromer_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    peaks = load_peaks()

    nu = np.linspace(NU[0], NU[1], M_CH)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
    src.simulate()
    src.sigma = np.full(M_CH, SIGMA)
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    assert (src.m, len(par.phi)) == (M_CH, N_PHI)
    plan = dv.DevicePlan.get(src.lambda2, src.l2_ref, par.phi, device)
    P, K = args.slab, args.iters
    lam0 = 40.0 * src.theo_noise
    noise = lam0 / (2.0 * K)

    s = torch.as_tensor(src.s, dtype=torch.float32, device=device)
    if WORKLOAD == "c4":
        made = [make_slab_c4(torch, src, P, 1235 + 17 * rank + i, device) for i in range(2)]
        slabs_dev = [dv.pack_planar(c, plan.m, plan.ld2m, device)[0] for c, _, _ in made]
        w_slabs, gal_slabs = [wt for _, wt, _ in made], [pg for _, _, pg in made]
        slabs_host = None
        del made
        k_slabs = [(wt / s[None, :]).sum(dim=1) for wt in w_slabs]
        lam_slabs = [40.0 / torch.sqrt(wt.sum(dim=1)) for wt in w_slabs]          # 40 theo_noise per pixel
        noise_slabs = [lt / (2.0 * K) for lt in lam_slabs]
        scratch = torch.empty_like(slabs_dev[0])
    else:
        slabs_c = [make_slab(torch, src, P, 1234 + 17 * rank + i, device) for i in range(2)]   # complex (m, P)
        slabs_dev = [dv.pack_planar(c, plan.m, plan.ld2m, device)[0] for c in slabs_c]          # fp32 [P, 2m]
        slabs_host = [c.cpu().pin_memory() for c in slabs_c]                                    # e2e inputs
        del slabs_c
    w = torch.as_tensor(src.w, dtype=torch.float32, device=device)
    k = torch.full((P,), float(src.k), dtype=torch.float32, device=device)
    lam_t = torch.full((P,), lam0, dtype=torch.float32, device=device)
    noise_t = torch.full((P,), noise, dtype=torch.float32, device=device)
    x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=device)
    norm_buf = torch.zeros(2, dtype=torch.float32, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        if WORKLOAD == "c4":
            j = i % 2
            scratch.copy_(slabs_dev[j])
            plan.derotate(scratch, gal_slabs[j])        # Dataset.subtract_galacticrm (base/dataset.py:377-381)
            out = plan.fista(scratch, w_slabs[j], s, k_slabs[j], x, lam_slabs[j], noise_slabs[j], K,
                             use_dirty_as_guess=True, want_delta=True, want_mma_blocks=True)
        else:
            out = plan.fista(slabs_dev[i % 2], w, s, k, x, lam_t, noise_t, K, use_dirty_as_guess=True, want_delta=True,
                             want_mma_blocks=True)
        # global convergence norm: the only collective on the path (mean |x_new - x_old| over all lines of sight)
        torch.sum(out["delta"], dim=0, out=norm_buf[0])
        norm_buf[1].fill_(float(P))
        if world > 1:
            dist.all_reduce(norm_buf)
        return out

    # ---- device-resident timing -------------------------------------------------------------------------
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.csr_launch_count(1)
    lib.csr_profile(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = device_step(i)
    e1.record()
    barrier()
    sampler.stop_flag = True
    launches = int(lib.csr_launch_count(0))
    ms = (np.zeros(12, dtype=np.float64), np.zeros(12, dtype=np.int64))
    lib.csr_profile(0)
    _lib.check(lib.csr_profile_collect(ms[0].ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                       ms[1].ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))), "csr_profile_collect")
    elapsed = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed.item())
    finite = bool(torch.isfinite(x).all())
    conv_norm = float(norm_buf[0] / norm_buf[1])
    value = world * P * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API with host buffers ----------------------------------------------
    # Streaming pipeline, as a cube run would do it: the device->host copy of step i's solution slab runs on a second
    # stream into one of two pinned buffers while step i+1 computes; every step's H2D and D2H stay inside the timed
    # region (the last copy is waited for before the clock stops).
    out_host = [torch.empty((2 * N_PHI, P), dtype=torch.float32).pin_memory() for _ in range(2)] if WORKLOAD == "c3" else None
    copy_stream = torch.cuda.Stream(device=device)
    copy_done = [None, None]

    def e2e_step(i):
        ds = cs.Dataset(lambda2=src.lambda2, data=slabs_host[i % 2], sigma=src.sigma, spectral_idx=1.0)
        dft = cs.NDFT1D(dataset=ds, parameter=par)
        f_dirty = dft.backward(ds.data)                      # H2D of the slab + dirty spectrum
        guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
        guess.data = f_dirty
        guess.complex_data_to_real()
        chi2 = cs.Chi2(dft_obj=dft, F_dirty=f_dirty)
        l1 = cs.L1(reg=lam0)
        opt = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                       maxiter=K, verbose=False)
        cost, X = opt.run()                                  # (returns after the D2H of the per-pixel cost)
        ready = torch.cuda.Event()
        ready.record()
        if copy_done[i % 2] is not None:
            copy_done[i % 2].synchronize()                   # the consumer has had this pinned buffer for a whole step
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            X.data.record_stream(copy_stream)
            out_host[i % 2].copy_(X.data, non_blocking=True)  # D2H of the solution cube slab
            copy_done[i % 2] = torch.cuda.Event()
            copy_done[i % 2].record(copy_stream)
        return cost

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if WORKLOAD == "c3":
        e2e_step(0)
        copy_stream.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            cost = e2e_step(i)
        copy_stream.synchronize()
        barrier()
        e2e_wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(e2e_wall, op=dist.ReduceOp.MAX)
        e2e_value = world * P * e2e_steps / float(e2e_wall.item())
    else:  # the optional workloads report the device-resident rate only
        e2e_value, cost = None, (out["cost"][:, 0] + lam_slabs[0] * 0.5 * out["cost"][:, 1]).cpu().numpy()
    h2d = P * M_CH * 8
    d2h = P * 2 * N_PHI * 4 + P * 8      # solution slab (fp32 planar) + per-pixel (chi2, l1) values

    # ---- dense reference leg: the same step with every exact shortcut switched off (bit-identical results) ------
    def profiled(steps):
        lib.csr_profile(1)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            o = device_step(i)
        t1.record()
        torch.cuda.synchronize()
        lib.csr_profile(0)
        pm = (np.zeros(12, dtype=np.float64), np.zeros(12, dtype=np.int64))
        _lib.check(lib.csr_profile_collect(pm[0].ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           pm[1].ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))), "csr_profile_collect")
        return t0.elapsed_time(t1), pm, o

    flops_launch = 8.0 * M_CH * N_PHI * P        # one transform = 8 m n algorithmic flops per line of sight
    peak = peaks["tf_sustained"]
    dense = None
    if not args.no_dense:
        _lib.set_option("zero_skip", 0)
        try:
            device_step(0)
            d_ms, d_pm, _ = profiled(1)
        finally:
            _lib.set_option("zero_skip", 1)
        d_gemm_ms, d_n = float(d_pm[0][:6].sum()), int(d_pm[1][:6].sum())
        d_ach = flops_launch * d_n / (d_gemm_ms * 1e-3) / 1e12
        dense = {"what": "same step, exact shortcuts off (zero_skip=0): every transform is a dense three-product GEMM",
                 "value": world * P / (d_ms * 1e-3), "unit": "LOS/s", "ms_per_step": d_ms,
                 "gemm_launches": d_n, "gemm_avg_ms": d_gemm_ms / d_n, "achieved": d_ach, "frac": d_ach / peak,
                 "executed_tflops": 3.0 * d_ach, "executed_frac": 3.0 * d_ach / peak,
                 "note": "3 fp16 MMAs per algorithmic product (hi/lo split, fp32-class accuracy): frac <= 1/3"}
        barrier()

    # ---- roofline ------------------------------------------------------------------------------------------------
    # A FISTA-adjoint transform is a cascade: fp8 screening launch over the tiles whose state is zero -> fp16
    # screening of the tiles that fail it -> three-product launch on what is left; together ONE transform.
    # `executed` is what the tensor cores really issued (in-kernel counters of 128 x 256 x 64 blocks).
    names = ["plain_forward", "plain_adjoint", "resid_forward", "resid_adjoint", "fista_forward", "fista_adjoint",
             "swt_synth_fused", "swt_analysis_update_fused", "screen_adjoint_fp16", "screen_adjoint_fp8"]
    gemm_kinds = [0, 1, 2, 3, 4, 5, 8, 9]
    gemm_ms = float(sum(ms[0][i] for i in gemm_kinds))
    transforms = int(sum(ms[1][i] for i in (0, 1, 2, 3, 4, 5)))
    path_ach = flops_launch * transforms / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    per_kind = {}
    for idx in gemm_kinds:
        if ms[1][idx]:
            per_kind[names[idx]] = {"launches": int(ms[1][idx]), "avg_ms": float(ms[0][idx] / ms[1][idx]),
                                    "share_of_step": float(ms[0][idx] / elapsed_ms)}
    block_flops = 2.0 * 128 * 256 * 64
    blocks = out["mma_blocks"].double().cpu().numpy()          # per step: [three-product, fp16 screen, fp8 screen, -]
    dominant = max(per_kind, key=lambda kname: per_kind[kname]["share_of_step"]) if per_kind else None
    # the fp8 screening kernel and the forward transform take about the same time per launch (0.30 ms each): report the
    # tensor-core kernel under `roofline` whenever it is within 10 % of the top share, the forward transform (bound by its
    # fused epilogue's HBM traffic) under `roofline.second_kernel`, so that the line does not flip between runs
    if dominant and "screen_adjoint_fp8" in per_kind and \
            per_kind["screen_adjoint_fp8"]["share_of_step"] >= 0.9 * per_kind[dominant]["share_of_step"]:
        dominant = "screen_adjoint_fp8"
    second = None
    if "resid_forward" in per_kind:
        f_ms = per_kind["resid_forward"]["avg_ms"]
        f_bytes = P * 2.0 * M_CH * (4 + 2 + 2 + 1)      # per element: b read (fp32), r_hi, r_lo (fp16) and r8 (fp8) written
        second = {"kernel": "dft_gemm_kernel<EPI_RESID>: forward transform with K-block skipping + fused residual epilogue",
                  "bound": "hbm", "achieved": f_bytes / (f_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                  "frac": f_bytes / (f_ms * 1e-3) / 1e9 / peaks["hbm"], "avg_ms": f_ms,
                  "share_of_step": per_kind["resid_forward"]["share_of_step"],
                  "note": "algorithmic bytes = 9 B per output element of the [P, 2m] residual; with ~9 of 249 K blocks active "
                          "the MMA work is negligible and the kernel is bound by the ~20 us fused epilogue of each 128 x 256 "
                          "tile (8 epilogue warps per SM, 32-byte row segments), not by HBM bandwidth: the next thing to fix"}
    if dominant == "screen_adjoint_fp8":
        k_ms = ms[0][9] / args.steps
        k_tf = blocks[2] * block_flops / (k_ms * 1e-3) / 1e12
        k_peak, k_src = 2.0 * peak, ("2 x MEASURED_PEAKS.json bf16_tflops_sustained (%s): no fp8 figure is measured on this "
                                     "pool; nominal fp8 dense = 2 x bf16" % peaks["source"])
        k_name = ("screen8_pair_kernel<3>: mirror-symmetric fp8 (e4m3) screening pass of the adjoint transform on CTA pairs "
                  "(+phi and -phi decided by one product with half the K), tcgen05 cta_group::2 kind::f8f6f4")
        k_note = ("achieved = fp8 flops the kernel issued (in-kernel counter) / its CUDA-event time; the kernel is bound by the "
                  "L2->SM operand fabric, not by the tensor pipe: ncu (profiles/r1_sym8_metrics.json) tensor pipe active 42 %, "
                  "2.54 GB of operands per launch = 7.4-8.5 TB/s, DRAM 42 MB; the symmetric formulation halves both the flops "
                  "and the operand bytes of the plain pair kernel (0.37 -> 0.30 ms per launch)")
    else:  # dense regime: the three-product kernels dominate
        k_ms = float(sum(ms[0][i] for i in (0, 1, 2, 3, 4, 5))) / args.steps
        k_tf = (blocks[0] + 3.0 * (ms[1][0] * ((P + 127) // 128) * ((2 * M_CH + 255) // 256) * ((2 * N_PHI + 63) // 64)
                                   + ms[1][1] * ((P + 127) // 128) * ((2 * N_PHI + 255) // 256) * ((2 * M_CH + 63) // 64))
                / args.steps) * block_flops / (k_ms * 1e-3) / 1e12
        k_peak, k_src = peak, "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peaks["source"]
        k_name = "dft_gemm_kernel (three fp16 products per algorithmic product)"
        k_note = "achieved = fp16 flops issued / CUDA-event time"
    adj_ms = float(ms[0][5] + ms[0][8] + ms[0][9])
    roofline = {"bound": "tensor", "kernel": k_name, "achieved": k_tf, "peak": k_peak, "unit": "TFLOP/s",
                "frac": k_tf / k_peak if k_peak else None, "traffic": ncu_traffic(P), "peak_source": k_src, "note": k_note,
                "kernel_share_of_step": per_kind[dominant]["share_of_step"] if dominant else None,
                "path": {"what": "all transform GEMM launches of the step: algorithmic 8 m n flops per line of sight per "
                                 "transform / their summed CUDA-event time. Exact sparsity shortcuts (forward K-block "
                                 "skipping, mirror-symmetric fp8 -> fp16 safe screening of the adjoint; bit-identical results, "
                                 "tests/test_gpu_parity.py::test_sparsity_shortcuts_are_bit_exact) let this exceed the "
                                 "dense bound of 1/3 of the tensor peak",
                         "algorithmic_tflops": path_ach, "frac_of_bf16_peak": path_ach / peak,
                         "gemm_share_of_step": gemm_ms / elapsed_ms if elapsed_ms > 0 else None,
                         "adjoint_transform_ms": adj_ms / ms[1][5] if ms[1][5] else None,
                         "executed_blocks_per_step": {"three_product": blocks[0], "fp16_screen": blocks[1], "fp8_screen": blocks[2]}},
                "second_kernel": second, "dense": dense, "per_kind": per_kind}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            v, wall, text = cpu_sample(args.cpu_iters, K, cores)
            cpu = {"value": v, "unit": "LOS/s", "cores": cores, "kind": "port", "sample": text, "wall_s": wall}
        line = {
            "metric": "lines-of-sight/sec (fixed-iter FISTA)", "value": value, "unit": "LOS/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (split-fp16 tensor-core "
            "products, fp32 accumulate)", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "LOS/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
            "algorithmic_tflops": value * 8.0 * M_CH * N_PHI * (2 * K + 2) / 1e12,
            "finite": finite, "convergence_norm": conv_norm, "final_cost_mean": float(np.mean(np.asarray(cost))),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS), help="c3 = the driver's line (default)")
    ap.add_argument("--slab", type=int, default=None, help="lines of sight per step per GPU (c3: 16384, c4: 8192)")
    ap.add_argument("--iters", type=int, default=100, help="FISTA iterations (fixed)")
    ap.add_argument("--cpu-iters", type=int, default=4, help="FISTA iterations in the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense reference leg of the roofline")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.slab is None:
        args.slab = WORKLOADS[args.workload]["slab"]
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
