"""``NUFFT1D`` -- the reference's fast operator, kept as API; arithmetic = the exact operator.

Reference: transformers/dfts/nufft.py:10-77 wraps pynufft 2023.1.1 (min-max interpolation, ``Jd=4``,
``Kd=Nd``), a third-party package that is not part of the reference tree; with ``solve=True`` its ``backward``
is a one-step CG *scaled* solve.  That arithmetic cannot be reproduced or pinned here (parity unpinned,
SURVEY.md section 8c).  This class keeps the constructor, the method names and the scalings the reference applies
around pynufft (``* n / k`` in backward, ``k s / (n w)`` in forward_normalized) but evaluates the operator
pynufft approximates -- the exact non-uniform DFT on the centred grid ``phi_j = cellsize (j - n/2)`` -- on the
GPU.  ``conv_size``, ``oversampling_factor``, ``solve``, ``solver`` and ``maxiter`` are accepted and ignored.
"""
from dataclasses import dataclass

import numpy as np

from .ft import FT


@dataclass(init=True, repr=True)
class NUFFT1D(FT):
    conv_size: int = None
    oversampling_factor: int = None
    normalize: bool = None
    solve: bool = None

    def __post_init__(self):
        super().__post_init__()
        self.conv_size = 4 if self.conv_size is None else self.conv_size
        self.oversampling_factor = 1 if self.oversampling_factor is None else self.oversampling_factor
        self.normalize = True if self.normalize is None else self.normalize
        self.solve = False if self.solve is None else self.solve
        self._configured = False

    def configure(self):
        self.device_plan()
        self._configured = True

    def forward(self, x):
        return self._forward(x, normalized=False)

    def forward_normalized(self, x):
        return self._forward(x, normalized=True)

    def backward(self, b, solver="cg", maxiter=1):
        # pynufft's adjoint carries 1/n; the reference then multiplies by n/k when ``normalize`` (nufft.py:68-69)
        n = len(self.parameter.phi)
        scale = 1.0 / np.asarray(self.dataset.k, dtype=np.float64) if self.normalize else 1.0 / n
        return self._adjoint(b, scale)

    def RMTF(self):
        ones = np.ones(np.shape(self.dataset.w), dtype=np.complex64)
        return self._adjoint(ones, 1.0 / np.asarray(self.dataset.k, dtype=np.float64))
