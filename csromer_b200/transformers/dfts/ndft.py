"""``NDFT1D`` (alias ``DFT1D``, the name the reference README uses) -- the exact irregular DFT.

Reference: transformers/dfts/ndft.py:10-43, which rebuilds the full m x n complex128 twiddle matrix with
``np.exp`` on every call and applies it with ``np.dot``.  Here the matrix is built once per sampling on the
GPU and applied as a tensor-core GEMM to all lines of sight at once; results are complex64 like the
reference's ``.astype(np.complex64)``.

``forward_normalized`` uses ``dataset.k`` -- the reference reads a non-existent ``self.k`` (ndft.py:24), the
intended value is the one transformers/dfts/nufft.py:55 uses.  Channels with ``w == 0`` return 0 (the reference
leaves them uninitialised).
"""
from dataclasses import dataclass

import numpy as np

from .ft import FT


@dataclass(init=True, repr=True)
class NDFT1D(FT):

    def __post_init__(self):
        super().__post_init__()

    def configure(self):
        self.device_plan()

    def forward(self, x):
        """P_i = sum_j x_j exp(+2i (lambda2_i - l2_ref) phi_j)"""
        return self._forward(x, normalized=False)

    def forward_normalized(self, x):
        """[k s_i / (n w_i)] sum_j x_j exp(+2i ...)"""
        return self._forward(x, normalized=True)

    def backward(self, b):
        """F_j = (1/k) sum_i (w_i / s_i) b_i exp(-2i (lambda2_i - l2_ref) phi_j)"""
        return self._adjoint(b, 1.0 / np.asarray(self.dataset.k, dtype=np.float64), remember=True)

    def RMTF(self, phi_x=0.0):
        """adjoint of the weights themselves (ndft.py:38-43); |R(0)| = 1 when s == 1"""
        ones = np.ones(np.shape(self.dataset.w), dtype=np.complex64)
        return self._adjoint(ones, 1.0 / np.asarray(self.dataset.k, dtype=np.float64))


DFT1D = NDFT1D
