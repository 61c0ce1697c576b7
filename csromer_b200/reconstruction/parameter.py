"""``Parameter`` -- the Faraday-depth-space container (phi grid + the vector being reconstructed).

Host-side mirror of the reference's ``reconstruction/parameter.py:26-137``.  ``data`` keeps phi (or the
coefficient index) on axis 0, so a batch is ``(n, P)`` complex, ``(2n, P)`` planar real or ``(ncoef, P)``
wavelet coefficients.  As in the reference, assigning ``data`` overwrites ``n`` with ``len(data)``
 -- operators therefore use ``len(parameter.phi)``, never ``parameter.n``.
``convolve`` (the restore step, parameter.py:139-169) runs on the GPU (``csrc/restore.cu``).
"""
import weakref

import numpy as np

from ..utils import complex_to_real, next_power_2, real_to_complex


def _is_complex(a):
    return np.iscomplexobj(a) if isinstance(a, np.ndarray) else bool(getattr(a, "is_complex", lambda: False)())


class Parameter:

    def __init__(self, phi=None, cellsize=None, data=None):
        self._data = None
        self._n = 0
        self.phi = phi
        self.cellsize = cellsize
        self.rmtf_fwhm = 0.0
        self.max_recovered_width = 0.0
        self.max_faraday_depth = 0.0
        if phi is not None:
            self._n = len(phi)
        self.data = data
        if phi is not None:  # phi wins over data for the initial n, like the reference constructor
            self._n = len(phi)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, val):
        self._data = val
        self._real_of = None
        if val is not None:
            self._n = len(val)

    @property
    def n(self):
        return self._n

    @n.setter
    def n(self, val):
        self._n = val

    def calculate_cellsize(self, dataset=None, oversampling=None, cellsize=None, set_size_pow_2=False, verbose=True):
        """phi grid from the lambda2 coverage: FWHM of the RMTF, largest recoverable scale, maximum depth,
        ``n`` rounded down to a multiple of 32 (or up to a power of two)."""
        if dataset is None:
            return
        l2 = np.asarray(dataset.lambda2)
        l2_min = np.min(l2[np.nonzero(l2)])
        l2_max = np.max(l2)
        fwhm = 2.0 * np.sqrt(3.0) / (l2_max - l2_min)
        widest = np.pi / l2_min
        phi_max = max(np.sqrt(3) / dataset.delta_l2_mean, fwhm * 10.0)
        self.rmtf_fwhm, self.max_recovered_width, self.max_faraday_depth = fwhm, widest, phi_max
        if verbose:
            print("FWHM of the main peak of the RMTF: {0:.3f} rad/m^2".format(fwhm))
            print("Maximum recovered width structure: {0:.3f} rad/m^2".format(widest))
            print("Maximum Faraday Depth to which one has more than 50% sensitivity: {0:.3f}".format(phi_max))
        if oversampling is None:
            oversampling = 8.0
        phi_r = min(fwhm, widest) / oversampling if cellsize is None else cellsize
        cells = int(np.int32(np.floor(2 * phi_max / phi_r)))
        n = next_power_2(cells) if set_size_pow_2 else int(cells - cells % 32)
        self.cellsize = 2 * phi_max / n
        self.phi = self.cellsize * np.arange(-(n / 2), (n / 2), 1)
        self.data = np.zeros_like(self.phi, dtype=np.complex64)
        self._n = n

    def complex_data_to_real(self):
        if not _is_complex(self._data):
            raise TypeError("Parameter data is not complex64")
        src = self._data
        self.data = complex_to_real(src)
        if hasattr(src, "_version"):   # torch: remember which complex tensor this planar copy was made from (FISTA._run_fused)
            self._real_of = (weakref.ref(src), src._version, weakref.ref(self._data), self._data._version)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_real_of"] = None   # (weak references: not picklable, and meaningless in another process)
        return state

    def complex_source(self):
        """the complex tensor ``data`` was made from by ``complex_data_to_real`` if neither has been modified since, else None"""
        r = getattr(self, "_real_of", None)
        if r is None:
            return None
        src, real = r[0](), r[2]()
        if src is None or real is None or real is not self._data or src._version != r[1] or real._version != r[3]:
            return None
        return src

    def real_data_to_complex(self):
        if _is_complex(self._data):
            raise ValueError("Parameter data is not real")
        self.data = real_to_complex(self._data)

    def convolve(self, x=None, rmtf_fwhm=None, verbose=True):
        """Restore beam: convolve the complex Faraday spectrum (``x`` or ``self.data``, ``(n,)`` or ``(n, *pix)``)
        with a Gaussian of FWHM ``rmtf_fwhm`` (default: the RMTF's), real and imaginary parts separately,
        ``mode="same"`` -- the reference's scipy/astropy pipeline, on the GPU for all lines of sight at once."""
        from .. import _device as dv
        if rmtf_fwhm is None:
            rmtf_fwhm = self.rmtf_fwhm
        taps, sigma_pix = dv.gaussian_beam_taps(rmtf_fwhm, self.cellsize)
        if verbose:
            fwhm_pix = int(np.round(rmtf_fwhm / self.cellsize))
            sigma = rmtf_fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))
            print("Convolving with Gaussian kernel where FWHM {0:2.3f} rad/m^2 - pixels {1}, sigma {2:2.3f} rad/m^2 - "
                  "pixels {3}".format(rmtf_fwhm, fwhm_pix, sigma, sigma_pix))
        src = self._data if x is None else x
        if not _is_complex(src):
            raise TypeError("convolve expects the complex Faraday spectrum")
        n = src.shape[0]
        ld = (2 * n + 7) // 8 * 8
        planar, pix = dv.pack_planar(src, n, ld)
        out = dv.convolve_phi(planar, n, taps)
        return dv.unpack_complex(out, n, pix, like_torch=dv.is_torch(src))
