"""Development checks run on the GPU box, one stage per process so that a hang in one stage cannot hide the
others:  python tools/gpu_check.py <stage>.  Not part of the product or of the test-suite."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import csromer_b200 as cs  # noqa: E402
from csromer_b200 import _device as dv  # noqa: E402
from oracle import restatement as ob  # noqa: E402


def setup(nch=200, P=5, seed=3, noise=True, per_pixel_w=False, flags=0.0):
    rng = np.random.RandomState(seed)
    nu = np.linspace(1.008e9, 2.031e9, nch)
    phi_gal = rng.uniform(-400, 400, size=P)
    amp = 0.0035 * rng.uniform(0.5, 1.5, size=P)
    src = cs.FaradayThinSource(nu=nu, s_nu=amp, phi_gal=phi_gal, spectral_idx=1.0)
    src.simulate()
    if noise:
        src.apply_noise(0.2 * 0.0035, random_state=rng)
    if per_pixel_w:
        sig = 0.0007 * rng.uniform(0.5, 2.0, size=(nch, P))
        src.sigma = sig
        if flags > 0:
            w = src.w.copy()
            w[rng.uniform(size=w.shape) < flags] = 0.0
            src.w = w
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    return src, par


def relerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def stage_transforms(nch, P):
    src, par = setup(nch, P)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    op = ob.ExactDFT(src.lambda2, par.phi, src.w, src.s, src.k)
    t0 = time.time()
    F = dft.backward(src.data)
    torch.cuda.synchronize()
    print("backward  m=%d n=%d P=%d  relerr=%.3e  (%.2fs)" % (src.m, len(par.phi), P, relerr(F, op.backward(src.data)), time.time() - t0))
    rng = np.random.RandomState(5)
    x = np.zeros((len(par.phi), P), dtype=np.complex128)
    for p in range(P):
        idx = rng.choice(len(par.phi), 9, replace=False)
        x[idx, p] = (rng.normal(size=9) + 1j * rng.normal(size=9)) * 1e-3
    print("forward            relerr=%.3e" % relerr(dft.forward(x), op.forward(x)))
    print("forward_normalized relerr=%.3e" % relerr(dft.forward_normalized(x), op.forward_normalized(x)))
    xd = rng.normal(size=(len(par.phi), P)) + 1j * rng.normal(size=(len(par.phi), P))
    print("forward (dense x)  relerr=%.3e" % relerr(dft.forward(xd), op.forward(xd)))
    print("rmtf               relerr=%.3e  |R|max=%.7f" % (relerr(dft.RMTF(), op.rmtf()), np.abs(dft.RMTF()).max()))


def stage_fista(nch, P, K, per_pixel_w=False, flags=0.0, wavelet=None):
    src, par = setup(nch, P, per_pixel_w=per_pixel_w, flags=flags)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    F = dft.backward(src.data)
    lam0 = 40.0 * np.asarray(src.theo_noise)
    noise = lam0 / (2.0 * K)
    wav = owav = None
    if wavelet:
        wav = cs.UndecimatedWavelet(wavelet_name=wavelet) if not wavelet.startswith("dwt:") else \
            cs.DiscreteWavelet(wavelet_name=wavelet[4:], mode="periodization")
        owav = ob.WaveletDict(wavelet if not wavelet.startswith("dwt:") else wavelet[4:],
                              kind="dwt" if wavelet.startswith("dwt:") else "swt")
    par.data = F
    par.complex_data_to_real()
    if wav is not None:
        par.data = wav.decompose(par.data)
    chi2 = cs.Chi2(dft_obj=dft, wavelet=wav)
    l1 = cs.L1(reg=lam0)
    opt = cs.FISTA(guess_param=par, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                   maxiter=K, verbose=False)
    t0 = time.time()
    cost, X = opt.run()
    torch.cuda.synchronize()
    dt = time.time() - t0
    op = ob.ExactDFT(src.lambda2, par.phi, src.w, src.s, src.k)
    x0 = ob.complex_to_real(op.backward(np.asarray(src.data)))
    if owav is not None:
        x0 = owav.decompose(x0)
    ref = ob.fista(op, np.asarray(src.data), x0, lam0, noise, K, wavelet=owav)
    good = np.asarray(src.w) > 0
    if good.ndim == 1:
        good = good[:, None] & np.ones((1, P), bool)
    print("fista m=%d n=%d P=%d K=%d ppw=%s flags=%.2f wav=%s: err(x)=%.3e err(model)=%.3e err(cost)=%.3e nnz=%d/%d  (%.2fs)" %
          (src.m, len(par.phi), P, K, per_pixel_w, flags, wavelet, relerr(X.data, ref["x"]),
           float(np.abs((src.model_data - ref["model"])[good]).max() / np.abs(ref["model"]).max()),
           relerr(cost, ref["cost"]), np.count_nonzero(X.data), np.count_nonzero(ref["x"]), dt))


def stage_wavelets():
    rng = np.random.RandomState(2)
    for name, kind, N, lvl, app, norm in [("db1", "swt", 64, None, False, False), ("coif2", "swt", 1568 * 2, None, False, False),
                                          ("sym2", "swt", 3136, 3, True, False), ("db4", "swt", 3136, None, False, True),
                                          ("db2", "dwt", 3136, None, False, False), ("coif2", "dwt", 4736, None, True, False),
                                          ("db1", "dwt", 250, None, False, False), ("bior2.2", "swt", 1024, 4, False, False)]:
        x = rng.normal(size=(N, 7))
        if kind == "swt":
            w = cs.UndecimatedWavelet(wavelet_name=name, wavelet_level=lvl, append_signal=app, norm=norm)
        else:
            w = cs.DiscreteWavelet(wavelet_name=name, wavelet_level=lvl, mode="periodization", append_signal=app)
        o = ob.WaveletDict(name, kind=kind, level=lvl, norm=norm, append_signal=app)
        c = w.decompose(x)
        co = o.decompose(x)
        r = w.reconstruct(c)
        ro = o.reconstruct(co)
        print("wavelet %-8s %s N=%d lvl=%s app=%s norm=%s: ncoef=%d/%d err(dec)=%.2e err(rec)=%.2e PR=%.2e" %
              (name, kind, N, lvl, app, norm, c.shape[0], co.shape[0], relerr(c, co), relerr(r, ro),
               relerr(r, x * (2 if app else 1))))


def stage_bench(nch, P, K):
    """raw device-resident timing of the fused loop at a bench-like size"""
    rng = np.random.RandomState(1)
    nu = np.linspace(1.008e9, 2.031e9, nch)
    src1 = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
    src1.simulate()
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src1, oversampling=8, verbose=False)
    plan = dv.DevicePlan.get(src1.lambda2, 0.0, par.phi)
    dev = plan.device
    m, n = plan.m, plan.n
    g = torch.Generator(device=dev).manual_seed(1234)
    data = 0.003 * torch.randn((P, plan.ld2m), generator=g, device=dev)
    w = torch.full((m,), 1.0 / 0.0007 ** 2, device=dev)
    s = torch.as_tensor(src1.s, dtype=torch.float32, device=dev)
    k = torch.full((P,), float((w / s).sum()), device=dev)
    x = torch.zeros((P, plan.ld2n), device=dev)
    theo = 1.0 / np.sqrt(m / 0.0007 ** 2)
    lam = torch.full((P,), 40 * theo, device=dev)
    nz = torch.full((P,), 40 * theo / (2 * K), device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.fista(data, w, s, k, x, lam, nz, K, use_dirty_as_guess=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        flops = 8.0 * m * n * (2 * K + 2) * P
        print("bench m=%d n=%d P=%d K=%d: %.1f ms  %.1f LOS/s  %.1f TF/s algorithmic (x3 executed = %.1f)" %
              (m, n, P, K, ms, P / ms * 1e3, flops / ms / 1e9, 3 * flops / ms / 1e9))
    print("finite:", bool(torch.isfinite(x).all()), "max|x|", float(x.abs().max()))


def stage_wav_bench(P=4096, K=10, name="coif2", kind="swt"):
    """wavelet-basis fused loop and the stand-alone transforms at C1 sampling: time and effective HBM GB/s"""
    nu = np.linspace(1.008e9, 2.031e9, 1000)
    src1 = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
    src1.simulate()
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src1, oversampling=8, verbose=False)
    plan = dv.DevicePlan.get(src1.lambda2, 0.0, par.phi)
    dev = plan.device
    m, n = plan.m, plan.n
    wav = cs.UndecimatedWavelet(wavelet_name=name) if kind == "swt" else cs.DiscreteWavelet(wavelet_name=name, mode="periodization")
    eng = wav.prepare(2 * n)
    g = torch.Generator(device=dev).manual_seed(1234)
    sig = torch.randn((P, eng.ldn), generator=g, device=dev)
    for label, fn, inp in (("decompose", eng.decompose, sig), ("reconstruct", eng.reconstruct, None)):
        if inp is None:
            inp = eng.decompose(sig)
        fn(inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(inp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        alg = P * (eng.ncoef + eng.N) * 4.0  # read the input once, write the output once
        print("%s %s %s N=%d ncoef=%d P=%d: %.3f ms  algorithmic %.0f GB/s" % (kind, name, label, eng.N, eng.ncoef, P, ms, alg / ms / 1e6))
    data = 0.003 * torch.randn((P, plan.ld2m), generator=g, device=dev)
    w = torch.full((m,), 1.0 / 0.0007 ** 2, device=dev)
    s = torch.as_tensor(np.ascontiguousarray(src1.s), dtype=torch.float32, device=dev)
    k = torch.full((P,), float((w / s).sum()), device=dev)
    x = torch.zeros((P, eng.ldc), device=dev)
    theo = 1.0 / np.sqrt(m / 0.0007 ** 2)
    lam = torch.full((P,), 40 * theo, device=dev)
    nz = torch.full((P,), 40 * theo / (2 * K), device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.fista(data, w, s, k, x, lam, nz, K, use_dirty_as_guess=True, wavelet=eng)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("fista %s %s P=%d K=%d: %.1f ms  %.2f ms/iteration  %.0f LOS/s at K=100 (extrapolated)" %
              (kind, name, P, K, ms, ms / K, P / (ms / K * 100) * 1e3))


if __name__ == "__main__":
    stage = sys.argv[1]
    torch.cuda.set_device(0)
    print("== stage", stage, torch.cuda.get_device_name(0), flush=True)
    if stage == "t_small":
        stage_transforms(200, 5)
    elif stage == "t_ragged":
        stage_transforms(300, 300)
    elif stage == "t_c1":
        stage_transforms(1000, 130)
    elif stage == "f_small":
        stage_fista(200, 5, 30)
    elif stage == "f_ragged":
        stage_fista(200, 261, 40)
    elif stage == "f_flags":
        stage_fista(200, 140, 40, per_pixel_w=True, flags=0.15)
    elif stage == "f_c1":
        stage_fista(1000, 64, 100)
    elif stage == "wavelets":
        stage_wavelets()
    elif stage == "f_swt":
        stage_fista(200, 40, 30, wavelet="coif2")
    elif stage == "f_dwt":
        stage_fista(200, 40, 30, wavelet="dwt:db2")
    elif stage == "wav_bench":
        stage_wav_bench()
        stage_wav_bench(name="db2", kind="dwt")
    elif stage == "bench_small":
        stage_bench(1000, 4096, 10)
    elif stage == "bench":
        stage_bench(1000, 32768, 20)
    else:
        raise SystemExit("unknown stage")
