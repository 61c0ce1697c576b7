"""``L1`` -- sparsity prior: value, smooth-approximation gradient and soft-threshold prox
(reference: objectivefunction/priors/l1.py:8-32).  Host implementation for the generic (non-fused) path; inside
the fused FISTA loop the soft threshold runs in the adjoint GEMM's epilogue (csrc/dft_gemm.cu, EPI_FISTA).
"""
from dataclasses import dataclass

import numpy as np

from ..fi import Fi

_TINY = np.finfo(np.float32).tiny


def approx_abs(x, epsilon):
    return np.sqrt(x * x + epsilon)


@dataclass(init=True, repr=True)
class L1(Fi):

    def __post_init__(self):
        super().__post_init__()

    def evaluate(self, x, epsilon=_TINY):
        return np.sum(approx_abs(x, epsilon), axis=0)

    def calculate_gradient(self, x, epsilon=_TINY):
        return x / approx_abs(x, epsilon)

    def calculate_prox(self, x, nu=0):
        return np.sign(x) * np.maximum(np.abs(x) - self.reg, 0.0)
