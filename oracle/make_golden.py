"""Generate tests/golden/*.npz by running the UNMODIFIED reference (Oracle A, oracle/ref_shim.py).

Run once in the build container (``python oracle/make_golden.py``); the fixtures are committed because
``/root/reference`` does not travel to the GPU box.  Every array below is produced by the reference's own
classes (Dataset / simulation / Parameter / NDFT1D(+k) / Chi2 / L1 / OFunction / FISTA); nothing from
``oracle/restatement.py`` or ``csromer_b200`` is involved.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_source(R, nu, kind, noise_seed, noise_frac=0.2):
    thin = R.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    thin.simulate()
    src = thin
    if kind == "mixed":
        thick = R.FaradayThickSource(nu=nu, s_nu=0.0035, phi_fg=140, phi_center=200, spectral_idx=1.0)
        thick.simulate()
        src = thin + thick
    if noise_seed is not None:
        src.apply_noise(noise_frac * 0.0035, random_state=np.random.RandomState(noise_seed))
    return src


def run_case(R, name, nu, kind, noise_seed, lam_mode, maxiter, oversampling=8):
    t0 = time.time()
    src = make_source(R, nu, kind, noise_seed)
    par = R.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=oversampling, verbose=False)
    dft = R.NDFT1DFixed(dataset=src, parameter=par)
    out = dict(
        nu=nu, lambda2=src.lambda2, data=np.asarray(src.data), w=src.w, sigma=src.sigma, s=src.s,
        k=np.float64(src.k), l2_ref=np.float64(src.l2_ref), spectral_idx=np.float64(src.spectral_idx),
        theo_noise=np.float64(np.nan if src.theo_noise is None else src.theo_noise),
        phi=par.phi, n=np.int64(len(par.phi)), cellsize=np.float64(par.cellsize),
        rmtf_fwhm=np.float64(par.rmtf_fwhm), max_recovered_width=np.float64(par.max_recovered_width),
        max_faraday_depth=np.float64(par.max_faraday_depth),
    )
    F_dirty = dft.backward(src.data)
    out["F_dirty"] = F_dirty
    out["rmtf"] = dft.RMTF()
    # a sparse test vector through forward / forward_normalized
    rng = np.random.RandomState(7)
    xt = np.zeros(len(par.phi), dtype=np.complex128)
    idx = rng.choice(len(par.phi), 12, replace=False)
    xt[idx] = (rng.normal(size=12) + 1j * rng.normal(size=12)) * 1e-3
    out["x_test"] = xt
    out["fwd_x_test"] = dft.forward(xt)
    out["fwdn_x_test"] = dft.forward_normalized(xt)

    if lam_mode is not None:
        if lam_mode == "fixed":
            lam0, noise = 1e-3, 1e-5
        else:  # SURVEY section 8d stable setting: lam0 = 40 theo_noise, lam ends at 20 theo_noise
            lam0 = 40.0 * src.theo_noise
            noise = lam0 / (2.0 * maxiter)
        par.data = F_dirty
        par.complex_data_to_real()
        chi2 = R.Chi2(dft_obj=dft, wavelet=None)
        l1 = R.L1(reg=lam0)
        F_obj = R.OFunction([chi2, l1])
        g_obj = R.OFunction([l1])
        opt = R.FISTA(guess_param=par, F_obj=F_obj, fx=chi2, gx=g_obj, noise=noise, maxiter=maxiter,
                      verbose=False)
        obj, X = opt.run()
        out.update(lam0=np.float64(lam0), noise=np.float64(noise), maxiter=np.int64(maxiter),
                   fista_obj=np.float64(obj), fista_x=np.asarray(X.data), lam_final=np.float64(l1.reg),
                   model_data=np.asarray(src.model_data), residual=np.asarray(src.residual))
    os.makedirs(OUT, exist_ok=True)
    # complex128 -> complex64 where the reference itself has rounded to complex64 keeps fixtures small
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s m=%d n=%d  %.1fs" % (name, src.m, len(par.phi), time.time() - t0))


def main():
    R = ref_shim.load()
    jvla1000 = np.linspace(1.008e9, 2.031e9, 1000)  # tests/conftest.py:13, README.md:90
    jvla200 = np.linspace(1.008e9, 2.031e9, 200)
    meerkat300 = np.linspace(0.9e9, 1.67e9, 300)  # tests/conftest.py:7 band
    run_case(R, "thin_nonoise_jvla1000", jvla1000, "thin", None, None, None)
    run_case(R, "thin_noise0_jvla1000_k100", jvla1000, "thin", 0, "stable", 100)
    run_case(R, "mixed_noise1_jvla1000_k30", jvla1000, "mixed", 1, "fixed", 30)
    run_case(R, "mixed_noise1_jvla200_k60", jvla200, "mixed", 1, "stable", 60)
    run_case(R, "thin_noise2_meerkat300_k40", meerkat300, "thin", 2, "stable", 40)


if __name__ == "__main__":
    main()
