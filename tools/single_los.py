"""BASELINE configs[0] and [1] timed end to end through the public API (one line of sight, host arrays):
[0] thin source, JVLA 1000 channels, DFT1D + L1 delta-basis FISTA with the reference's DEFAULT schedule
    (lambda = sqrt(m + 2 sqrt(m)) sqrt(2) mean(sigma), noise = theo_noise, 1458 iterations; the survey measured 890 s
    per line of sight for the reference on 8 vCPU);
[1] thin + thick mixed source, undecimated coif2 basis, 20 % channels flagged (w = 0), NUFFT1D API."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csromer_b200 as cs  # noqa: E402

nu = np.linspace(1.008e9, 2.031e9, 1000)


def config0():
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=np.random.RandomState(0))
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    dft = cs.DFT1D(dataset=src, parameter=par)
    nufft = cs.NUFFT1D(dataset=src, parameter=par, solve=True)
    t0 = time.perf_counter()
    F_dirty = dft.backward(src.data)
    par.data = F_dirty
    par.complex_data_to_real()
    lam = np.sqrt(src.m + 2 * np.sqrt(src.m)) * np.sqrt(2) * np.mean(src.sigma)
    chi2 = cs.Chi2(dft_obj=nufft, wavelet=None)
    l1 = cs.L1(reg=lam)
    opt = cs.FISTA(guess_param=par, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=src.theo_noise,
                   verbose=False)
    obj, X = opt.run()
    X.real_data_to_complex()
    rms = float(np.sqrt(np.mean(np.abs(src.residual) ** 2)))
    dt = time.perf_counter() - t0
    nz = int(np.count_nonzero(X.data))
    print("config0: lambda0=%.6g iterations=%d objective=%.4f nonzero cells=%d argmax phi=%.3f max|X|=%.5f final lambda=%.4g "
          "rms(residual)=%.5g  wall %.3f s" % (lam, int(np.floor(lam / src.theo_noise)), obj, nz,
                                              par.phi[np.argmax(np.abs(X.data))], np.abs(X.data).max(), l1.reg, rms, dt))
    print("         reference (SURVEY 8c): objective 644.0708 nonzero 2719 argmax -201.958 max|X| 0.0192148 final lambda 5.519e-06 "
          "rms 0.00079447 (its last ~15 iterations run in the unit step's unstable region)")
    return dt


def config1():
    rng = np.random.RandomState(1)
    thin = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    thick = cs.FaradayThickSource(nu=nu, s_nu=0.0035, phi_fg=140, phi_center=200, spectral_idx=1.0)
    thin.simulate()
    thick.simulate()
    src = thin + thick
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    w = src.w.copy()
    for start in rng.randint(0, 950, size=6):
        w[start:start + 34] = 0.0
    src.w = w
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    dft = cs.DFT1D(dataset=src, parameter=par)
    nufft = cs.NUFFT1D(dataset=src, parameter=par, solve=True)
    wav = cs.UndecimatedWavelet(wavelet_name="coif2")
    t0 = time.perf_counter()
    par.data = dft.backward(src.data)
    par.complex_data_to_real()
    par.data = wav.decompose(par.data)
    K = 100
    lam = 40 * src.theo_noise
    chi2 = cs.Chi2(dft_obj=nufft, wavelet=wav)
    l1 = cs.L1(reg=lam)
    opt = cs.FISTA(guess_param=par, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=lam / (2 * K), maxiter=K,
                   verbose=False)
    obj, X = opt.run()
    X.data = wav.reconstruct(X.data)
    X.real_data_to_complex()
    dt = time.perf_counter() - t0
    print("config1: flagged %.0f%% K=%d objective=%.4f max|X|=%.5f  wall %.3f s" % (100 * np.mean(src.w == 0), K, obj, np.abs(X.data).max(), dt))
    return dt


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for rep in range(2):
        config0()
        config1()
