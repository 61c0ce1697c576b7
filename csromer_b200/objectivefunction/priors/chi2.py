"""``Chi2`` -- data fidelity 0.5 sum w |d - P_hat(x)|^2 and its FISTA gradient
(reference: objectivefunction/priors/chi2.py:10-62).

Every method keeps the reference's side effects: ``evaluate`` and ``calculate_gradient_fista`` assign
``dft_obj.dataset.model_data`` (which refreshes ``dataset.residual``), and the constructor caches
``F_dirty = dft_obj.backward(dataset.data)``.  The transforms and wavelet steps run on the GPU through
``dft_obj`` / ``wavelet``.  ``FISTA.run`` recognises a ``Chi2`` + ``L1`` pair and replaces the Python loop over
these methods by the fused device loop (``csr_fista``).
"""
from dataclasses import dataclass

import numpy as np

from ...transformers.dfts import FT
from ...utils.utilities import complex_to_real, real_to_complex
from ..fi import Fi


@dataclass(init=True, repr=True)
class Chi2(Fi):
    dft_obj: FT = None
    F_dirty: np.ndarray = None

    def __post_init__(self):
        super().__post_init__()
        if self.dft_obj is not None and self.F_dirty is None:
            self.F_dirty = self.dft_obj.backward(self.dft_obj.dataset.data)

    def _phi_space(self, x):
        xs = self.wavelet.reconstruct(x.copy()) if self.wavelet is not None else x.copy()
        return real_to_complex(xs) * self.norm_factor

    def evaluate(self, x):
        ds = self.dft_obj.dataset
        ds.model_data = self.dft_obj.forward_normalized(self._phi_space(x))
        res = ds.residual
        w = ds.w if np.ndim(ds.w) == np.ndim(res) else np.reshape(ds.w, (-1,) + (1,) * (np.ndim(res) - 1))
        return 0.5 * np.sum(w * (res.real ** 2 + res.imag ** 2), axis=0)

    def calculate_gradient(self, x):
        return complex_to_real(self._phi_space(x) - self.F_dirty)

    def calculate_gradient_fista(self, x):
        ds = self.dft_obj.dataset
        ds.model_data = self.dft_obj.forward_normalized(self._phi_space(x))
        grad = complex_to_real(self.dft_obj.backward(-ds.residual))
        if self.wavelet is not None:
            grad = self.wavelet.decompose(grad)
        return grad

    def calculate_prox(self, x, nu=0):
        return (self.reg * complex_to_real(self.F_dirty) + nu) / (self.reg + 1.0)
