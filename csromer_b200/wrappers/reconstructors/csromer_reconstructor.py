"""``CSROMERReconstructorWrapper`` -- the reference's one-call reconstruction of a line of sight (or, here, of a
batch of them) on top of the GPU hot path (reference: wrappers/reconstructors/csromer_reconstructor.py:15-241).

Same attributes, defaults and sequence as the reference: optional flagging, dirty Faraday spectrum and its peak
statistics, lambda from the mean sigma (``sqrt(m + 2 sqrt(m)) sqrt(2) mean(sigma)``, doubled with a wavelet),
``FISTA(noise=theo_noise)`` with ``Chi2(dft_obj=NUFFT1D(solve=True))`` and ``L1``, model / residual / restored spectra,
RM at the peaks with the 3-point parabola, ``sigma_phi = FWHM / (2 SNR)`` and the second moment of the model.

Not reproducible bit for bit: the Faraday-depth noise uses ``astropy.stats.sigma_clipped_stats(mask=|phi| >
max_depth * threshold, sigma=0.3, cenfunc='mean', stdfunc='mad_std')`` (csromer_reconstructor.py:86-102) and astropy is
not available here; ``sigma_clipped_std`` below restates its documented algorithm (parity unpinned).  With the
reference's default ``threshold=0`` that mask leaves only the phi = 0 cell, so the reference's own noise is 0 there;
the same happens here.  ``step`` is this package's extension (the reference's unit step is unstable at the end of the
default lambda schedule with the exact operator): ``"auto"`` = 1 / lambda_max.
"""
from dataclasses import dataclass, field

import numpy as np

from ...cube import estimate_lipschitz
from ...dictionaries import Wavelet
from ...objectivefunction import L1, Chi2, OFunction
from ...optimization import FISTA
from ...reconstruction.parameter import Parameter
from ...transformers.dfts import NDFT1D, NUFFT1D
from ...transformers.flaggers.flagger import Flagger
from .faraday_reconstructor import FaradayReconstructorWrapper


def sigma_clipped_std(values, mask=None, sigma=3.0, maxiters=5, cenfunc="mean", stdfunc="mad_std"):
    """standard deviation of ``values[~mask]`` after iterative clipping of |x - center| > sigma * spread
    (astropy.stats.sigma_clipped_stats as documented: the third returned value)"""
    x = np.asarray(values, dtype=np.float64)
    if mask is not None:
        x = x[~np.asarray(mask, dtype=bool)]
    x = x[np.isfinite(x)]
    for _ in range(maxiters):
        if x.size == 0:
            break
        center = np.mean(x) if cenfunc == "mean" else np.median(x)
        spread = 1.482602218505602 * np.median(np.abs(x - np.median(x))) if stdfunc == "mad_std" else np.std(x)
        keep = np.abs(x - center) <= sigma * spread
        if keep.all():
            break
        x = x[keep]
    return float(np.std(x)) if x.size else float("nan")


@dataclass(init=True, repr=True)
class CSROMERReconstructorWrapper(FaradayReconstructorWrapper):
    parameter: Parameter = field(init=False)
    flagger: Flagger = None
    dft: NDFT1D = field(init=False)
    nufft: NUFFT1D = field(init=False)
    wavelet: Wavelet = None
    cellsize: float = None
    oversampling: float = None
    lambda_l_norm: float = None
    calculate_l2_zero: bool = None
    step: object = 1.0       # extension: gradient step ("auto" = 1 / lambda_max); 1.0 = the reference's unit step
    verbose: bool = False

    def __post_init__(self):
        if self.oversampling is None:
            self.oversampling = 7.0
        if self.calculate_l2_zero is None:
            self.calculate_l2_zero = False
        if self.calculate_l2_zero:
            print("Calculating l2_0")
            self.dataset.l2_ref = self.dataset.calculate_l2ref()
        self.parameter = Parameter()
        self.config_fd_space(self.cellsize, self.oversampling)
        self.config_fourier_transforms()

    # ---- static helpers (csromer_reconstructor.py:59-108); arrays may carry trailing pixel axes ---------------------
    @staticmethod
    def estimate_peak_quadratic_interpolation(fd_signal, cellsize):
        mag = np.abs(np.asarray(fd_signal))
        n = mag.shape[0]
        i0 = np.argmax(mag, axis=0)
        take = lambda idx: np.take_along_axis(mag, np.expand_dims(idx % n, 0), axis=0)[0] if mag.ndim > 1 else mag[idx % n]
        f0, fm, fp = take(i0), take(i0 - 1), take(i0 + 1)
        pos = (fp - fm) / (4 * f0 - 2 * fm - 2 * fp)
        peak = f0 - 0.25 * (fm - fp) * pos
        return (i0 + pos - n / 2) * cellsize, peak

    @staticmethod
    def calculate_ricean_peak(peak, noise):
        return np.sqrt(peak ** 2 - (2.3 * noise ** 2))

    @staticmethod
    def calculate_fd_signal_noise(fd_signal, phi, max_fd_depth, threshold=0., sigma=0.3, cenfunc="mean", stdfunc="mad_std"):
        mask = np.abs(phi) > max_fd_depth * threshold
        fd = np.asarray(fd_signal)
        if fd.ndim == 1:
            return 0.5 * (sigma_clipped_std(fd.real, mask, sigma, 5, cenfunc, stdfunc) +
                          sigma_clipped_std(fd.imag, mask, sigma, 5, cenfunc, stdfunc))
        flat = fd.reshape(fd.shape[0], -1)
        out = np.array([0.5 * (sigma_clipped_std(flat[:, p].real, mask, sigma, 5, cenfunc, stdfunc) +
                               sigma_clipped_std(flat[:, p].imag, mask, sigma, 5, cenfunc, stdfunc))
                        for p in range(flat.shape[1])])
        return out.reshape(fd.shape[1:])

    @staticmethod
    def calculate_sigma_phi_peak(rmtf_fwhm, fd_peak, fd_signal_noise):
        return rmtf_fwhm / (2. * fd_peak / fd_signal_noise)

    # ---- configuration ---------------------------------------------------------------------------------------------
    def flag_dataset(self, flagger=None):
        return (self.flagger if flagger is None else flagger).run()

    def config_fd_space(self, cellsize=None, oversampling=None):
        if cellsize is not None:
            self.parameter.calculate_cellsize(dataset=self.dataset, cellsize=cellsize, verbose=self.verbose)
        elif oversampling is not None:
            self.parameter.calculate_cellsize(dataset=self.dataset, oversampling=oversampling, verbose=self.verbose)
        else:
            raise ValueError("Either cellsize or oversampling cannot be Nonetype values")

    def config_fourier_transforms(self):
        self.dft = NDFT1D(dataset=self.dataset, parameter=self.parameter)
        self.nufft = NUFFT1D(dataset=self.dataset, parameter=self.parameter, solve=True)

    def get_dirty_faraday_depth(self):
        return self.dft.backward(self.dataset.data)

    def get_rmtf(self):
        return self.dft.RMTF()

    def get_rm(self, fd_data):
        return self.parameter.phi[np.argmax(np.abs(fd_data), axis=0)]

    # ---- the reconstruction (csromer_reconstructor.py:138-228) -----------------------------------------------------
    def reconstruct(self):
        if self.flagger:
            self.flag_dataset()
            self.config_fd_space(self.cellsize, self.oversampling)   # the sampling may have changed
            self.config_fourier_transforms()
        ds, par = self.dataset, self.parameter
        fd_dirty = np.asarray(self.get_dirty_faraday_depth())
        dirty_noise = self.calculate_fd_signal_noise(fd_dirty, par.phi, par.max_faraday_depth)
        self.fd_dirty = fd_dirty
        par.data = fd_dirty
        self.rm_dirty = self.get_rm(fd_dirty)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.rm_dirty_error = self.calculate_sigma_phi_peak(par.rmtf_fwhm, np.max(np.abs(fd_dirty), axis=0), dirty_noise)
            (self.rm_dirty_quadratic_interpolation,
             self.dirty_peak_quadratic_interpolation) = self.estimate_peak_quadratic_interpolation(fd_dirty, par.cellsize)
            self.rm_dirty_quadratic_interpolation_error = self.calculate_sigma_phi_peak(
                par.rmtf_fwhm, self.dirty_peak_quadratic_interpolation, dirty_noise)
        par.complex_data_to_real()
        if self.lambda_l_norm is None:
            factor = 2.0 * np.sqrt(2) if self.wavelet is not None else np.sqrt(2)
            self.lambda_l_norm = np.sqrt(ds.m + 2 * np.sqrt(ds.m)) * factor * np.mean(ds.sigma)
        if self.wavelet is not None:
            par.data = self.wavelet.decompose(par.data)
        chi2 = Chi2(dft_obj=self.nufft, wavelet=self.wavelet)
        l1 = L1(reg=self.lambda_l_norm)
        opt_noise = 2.0 * ds.theo_noise if self.wavelet is not None else ds.theo_noise
        step = 1.0 / estimate_lipschitz(self.dft) if self.step == "auto" else float(self.step)
        opt = FISTA(guess_param=par, F_obj=OFunction([chi2, l1]), fx=chi2, gx=OFunction([l1]), noise=opt_noise,
                    verbose=self.verbose, step=step)
        self.objective_value, X = opt.run()
        if self.wavelet is not None:
            self.coefficients = X.data
            X.data = self.wavelet.reconstruct(X.data)
        X.real_data_to_complex()
        self.fd_model = np.asarray(X.data)
        self.rm_model = self.get_rm(self.fd_model)
        self.second_moment = self.calculate_second_moment()
        self.fd_residual = np.asarray(self.dft.backward(ds.data - ds.model_data))
        self.fd_restored = np.asarray(X.convolve(verbose=self.verbose)) + self.fd_residual
        restored_noise = self.calculate_fd_signal_noise(self.fd_restored, par.phi, par.max_faraday_depth)
        self.rm_restored = self.get_rm(self.fd_restored)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.rm_restored_error = self.calculate_sigma_phi_peak(par.rmtf_fwhm, np.max(np.abs(self.fd_restored), axis=0),
                                                                   restored_noise)
            (self.rm_restored_quadratic_interpolation,
             self.restored_peak_quadratic_interpolation) = self.estimate_peak_quadratic_interpolation(self.fd_restored,
                                                                                                      par.cellsize)
            self.rm_restored_quadratic_interpolation_error = self.calculate_sigma_phi_peak(
                par.rmtf_fwhm, self.restored_peak_quadratic_interpolation, restored_noise)

    def calculate_second_moment(self):
        """intensity-weighted variance of phi over the non-zero cells of the model (csromer_reconstructor.py:230-241)"""
        mag = np.abs(self.fd_model)
        phi = self.parameter.phi.reshape((-1,) + (1,) * (mag.ndim - 1))
        k = np.sum(mag, axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            first = np.sum(phi * mag, axis=0) / k
            return np.sum(mag * (phi - first) ** 2, axis=0) / k
