"""``reconstruct_cube`` -- the README's cube recipe (``README.md:282-377``), batched.

The reference forks one Python process per pixel; each builds a Dataset / Parameter / DFT1D / NUFFT1D, estimates
the noise from the edges of the dirty Faraday spectrum, sets ``lambda = sqrt(2m + sqrt(4m)) * noise``, runs FISTA
(undecimated coif2 basis by default, ``floor(lambda / noise)`` iterations), and stores four planes per pixel:
dirty, model, restored (= beam * model + residual spectrum) and residual.  Here a slab of pixels goes through the
same steps as a handful of batched GPU calls; slabs are walked sequentially so a cube larger than HBM streams
through (``out`` may be a ``numpy.memmap`` like the README suggests).

Deviation that matters: the recipe's schedule lets lambda decay to ~0, where the unit gradient step is unstable with
the exact operator (the reference relies on pynufft's scaled solve there).  ``step`` therefore
defaults to ``"auto"`` = 1 / lambda_max((1/n) E^H E), estimated once by power iteration; ``step=1.0`` reproduces the
reference's unit step.
"""
import numpy as np
import torch

from . import _device as dv
from .base.dataset import Dataset
from .dictionaries import UndecimatedWavelet
from .objectivefunction import L1, Chi2, OFunction
from .optimization import FISTA
from .reconstruction.parameter import Parameter
from .transformers.dfts import NDFT1D, NUFFT1D


def estimate_lipschitz(dft, iterations=30, seed=0):
    """largest eigenvalue of (1/n) E^H E for the operator's sampling (pixel independent: weights cancel)"""
    plan = dft.device_plan()
    g = torch.Generator(device=plan.device).manual_seed(seed)
    v = torch.randn((1, plan.ld2n), generator=g, device=plan.device)
    v[:, 2 * plan.n:] = 0
    lam = 1.0
    for _ in range(iterations):
        v = v / v.norm()
        w = plan.adjoint(plan.forward(v)) / plan.n
        lam = float((w * v).sum())
        v = w
    return lam


def reconstruct_cube(data, nu, sigma=None, spectral_idx=None, eta=1.0, use_wavelet=True, wavelet_name="coif2",
                     oversampling=8, step="auto", slab=8192, out=None, maxiter=None, verbose=False):
    """``data`` complex ``(m, Y, X)`` or ``(m, P)`` (host array / memmap / tensor), ``nu`` (m,), ``sigma`` (m,) or None.
    Returns ``F`` complex64 ``(4, n, *pix)``: dirty, model, restored, residual Faraday spectra (written into ``out``
    if given), plus a dict with the per-pixel noise, lambda and peak statistics."""
    m = data.shape[0]
    pix = tuple(data.shape[1:])
    flat = data.reshape(m, -1)
    P_total = flat.shape[1]
    probe = Dataset(nu=nu, sigma=sigma, spectral_idx=spectral_idx)
    par = Parameter()
    par.calculate_cellsize(dataset=probe, oversampling=oversampling, verbose=verbose)
    n = len(par.phi)
    F = out if out is not None else np.zeros((4, n) + pix, dtype=np.complex64)
    F_flat = F.reshape(4, n, -1)
    edge = np.abs(par.phi) > par.max_faraday_depth / 1.5
    taps, _ = dv.gaussian_beam_taps(par.rmtf_fwhm, par.cellsize)
    info = dict(noise=np.zeros(P_total, np.float32), lambda_l1=np.zeros(P_total, np.float32),
                rm_model=np.zeros(P_total), rm_restored=np.zeros(P_total), peak_restored=np.zeros(P_total, np.float32),
                rm_restored_interpolated=np.zeros(P_total), step=None, iterations=None, phi=par.phi)
    edge_dev = None
    step_value = None
    info["blanked_samples"] = 0
    info["dead_pixels"] = np.zeros(P_total, dtype=bool)
    for lo in range(0, P_total, slab):
        hi = min(P_total, lo + slab)
        block = np.ascontiguousarray(flat[:, lo:hi]) if isinstance(flat, np.ndarray) else flat[:, lo:hi]
        ds = Dataset(nu=nu, sigma=sigma, data=block, spectral_idx=spectral_idx)
        # Blanked samples (NaN / inf: real Q/U cubes almost always carry some).  The reference's recipe never feeds such
        # pixels in (it selects pixels with utils.make_mask_faraday, README.md:298-307); here a non-finite SAMPLE becomes
        # 0 with weight 0 -- a flagged channel of that pixel only -- and a pixel without a single finite sample ("dead")
        # keeps the shared weights on all-zero data: its dirty spectrum, noise and lambda are 0, the loop freezes it
        # (noise >= lambda, fista.py:74-77) and its output planes are written as NaN.
        finite = torch.isfinite(block) if dv.is_torch(block) else np.isfinite(block)
        dead = None
        if not bool(finite.all()):
            if dv.is_torch(block):
                finite = finite.cpu().numpy()
                block = torch.where(torch.as_tensor(finite, device=block.device), block, torch.zeros_like(block))
            else:
                block = np.where(finite, block, 0)
            dead = ~finite.any(axis=0)
            info["blanked_samples"] += int(finite.size - finite.sum())
            info["dead_pixels"][lo:hi] = dead
            w_pp = np.asarray(probe.w, dtype=np.float64)[:, None] * (finite | dead[None, :])
            ds = Dataset(nu=nu, sigma=sigma, data=block, spectral_idx=spectral_idx)
            ds.w = w_pp                                                              # per-pixel flags: refreshes k, theo_noise
        dft = NDFT1D(dataset=ds, parameter=par)
        nufft = NUFFT1D(dataset=ds, parameter=par, solve=True)
        plan = dft.device_plan()
        if step_value is None:
            step_value = 1.0 / estimate_lipschitz(dft) if step == "auto" else float(step)
            edge_dev = torch.as_tensor(edge.astype(np.uint8)).to(plan.device)
        d_planar, _ = dv.pack_planar(ds.data, m, plan.ld2m, plan.device)
        w = np.asarray(ds.w)
        chan = dft._channel_tensor(w / (ds.s if w.ndim == 1 else np.asarray(ds.s)[:, None]), plan)   # [m] or [P, m]
        row = dv.per_pixel_vector(1.0 / np.asarray(ds.k), hi - lo, plan.device, "k")
        dirty = plan.adjoint(d_planar, chan, row)                                  # F_dirty = dft.backward(data)
        noise = dv.edge_noise(dirty, n, edge_dev, eta)                              # README.md:336-338
        lam = float(np.sqrt(2 * m + np.sqrt(4 * m))) * noise                        # README.md:350
        noise_h, lam_h = noise.cpu().numpy().astype(np.float64), lam.cpu().numpy().astype(np.float64)
        wav = UndecimatedWavelet(wavelet_name=wavelet_name) if use_wavelet else None
        guess = Parameter(phi=par.phi, cellsize=par.cellsize)
        x0 = dirty[:, :2 * n].t().contiguous()                                      # (2n, P) real planar, on the device
        guess.data = wav.decompose(x0) if wav is not None else x0
        chi2 = Chi2(dft_obj=nufft, wavelet=wav, F_dirty=dv.unpack_complex(dirty, n, (hi - lo,), like_torch=True))
        l1 = L1(reg=lam_h)
        opt = FISTA(guess_param=guess, F_obj=OFunction([chi2, l1]), fx=chi2, gx=OFunction([l1]), noise=noise_h,
                    maxiter=maxiter, verbose=False, step=step_value)
        _, X = opt.run()
        info["iterations"] = int(np.floor(np.sqrt(2 * m + np.sqrt(4 * m)))) if maxiter is None else maxiter
        model_real = wav.reconstruct(X.data) if wav is not None else X.data        # (2n, P) device
        model_planar = torch.zeros((hi - lo, plan.ld2n), dtype=torch.float32, device=plan.device)
        model_planar[:, :2 * n] = model_real.t()
        resid_planar = opt.device_outputs["residual"]                               # dataset.residual of the final evaluate
        f_resid = plan.adjoint(resid_planar, chan, row)                             # dft.backward(dataset.residual)
        restored = dv.convolve_phi(model_planar, n, taps, add=f_resid)             # X.convolve() + F_residual
        for plane, t in enumerate((dirty, model_planar, restored, f_resid)):
            F_flat[plane, :, lo:hi] = dv.unpack_complex(t, n, (hi - lo,))
        if dead is not None and dead.any():
            F_flat[:, :, lo + np.nonzero(dead)[0]] = np.nan
        idx_m, _ = dv.peak_stats(model_planar, n)
        idx_r, st_r = dv.peak_stats(restored, n)
        idx_m, idx_r, st_r = idx_m.cpu().numpy(), idx_r.cpu().numpy(), st_r.cpu().numpy()
        info["noise"][lo:hi], info["lambda_l1"][lo:hi] = noise_h, lam_h
        info["rm_model"][lo:hi] = par.phi[idx_m]
        info["rm_restored"][lo:hi] = par.phi[idx_r]
        info["peak_restored"][lo:hi] = st_r[:, 2]
        info["rm_restored_interpolated"][lo:hi] = (idx_r + st_r[:, 1] - n / 2) * par.cellsize
    info["step"] = step_value
    for key in ("noise", "lambda_l1", "rm_model", "rm_restored", "peak_restored", "rm_restored_interpolated", "dead_pixels"):
        info[key] = info[key].reshape(pix) if pix else info[key]
    return F, info
