"""``Wavelet`` -- base of the analysis/synthesis dictionaries (reference: dictionaries/wavelet.py:8-41).

The reference delegates to PyWavelets on the host, one line of sight at a time.  Here the transforms run on the
GPU for every line of sight of the batch (``csrc/wavelet.cu``); this class keeps the reference's constructor
fields and its statefulness: ``decompose`` records the signal length and band layout, ``reconstruct`` needs a
previous ``decompose`` (or an explicit ``prepare(length)``).

Arrays keep the sample / coefficient index on axis 0: ``(N,)`` or ``(N, *pix)`` in, ``(ncoeffs, *pix)`` out.
``decompose`` acts on the REAL planar vector ``[Re | Im]`` as one signal, as objectivefunction/priors/chi2.py:44-55
uses it; the ``*_complex`` variants transform real and imaginary parts separately.
"""
from abc import ABCMeta, abstractmethod
from dataclasses import dataclass, field

import numpy as np
import torch

from .. import _device as dv
from . import filters


@dataclass(init=True, repr=True)
class Wavelet(metaclass=ABCMeta):
    wavelet_name: str = None
    wavelet_level: int = None
    mode: str = None
    append_signal: bool = None
    ncoeffs: int = field(init=False, default=0)
    n: int = field(init=False, default=0)
    wavelet: object = field(init=False, default=None)  # the filter bank (dec_lo, dec_hi, rec_lo, rec_hi)
    coeff_slices: object = field(init=False, default=None)
    filter_bank: object = field(default=None, repr=False, kw_only=True)

    _kind = None  # 0 undecimated, 1 decimated

    def __post_init__(self):
        if not isinstance(self.wavelet_name, str):
            raise TypeError("The wavelet name is not a string")
        self.coeff_shapes = None
        self._engines = {}

    # ---- to be provided by the concrete dictionaries -----------------------------------------------------
    @abstractmethod
    def calculate_max_level(self, x):
        return

    def _level_for(self, length):
        raise NotImplementedError

    def _norm_flag(self):
        return False

    # ---- device engine --------------------------------------------------------------------------------------
    def prepare(self, length):
        """Fix the signal length (what ``decompose`` would record) and return the device engine for it."""
        length = int(length)
        eng = self._engines.get(length)
        if eng is None:
            eng = dv.DeviceWavelet(self._kind, self.wavelet, self._level_for(length), length,
                                   bool(self.append_signal), self._norm_flag())
            self._engines[length] = eng
        self.n = length
        self.ncoeffs = eng.ncoef
        offs = np.concatenate([[0], np.cumsum(eng.sizes)])
        self.coeff_slices = [slice(int(offs[i]), int(offs[i + 1])) for i in range(len(eng.sizes))]
        self.coeff_shapes = [(int(s),) for s in eng.sizes]
        return eng

    def device_engine(self):
        if not self.n:
            raise RuntimeError("decompose() must be called before reconstruct()")
        return self.prepare(self.n)

    def _decompose_real(self, x):
        like_torch = dv.is_torch(x)
        length = x.shape[0]
        eng = self.prepare(length)
        sig, pix = dv.pack_real(x, length, eng.ldn, eng.device)
        coef = eng.decompose(sig)
        out = dv.unpack_real(coef, eng.ncoef, pix, like_torch)
        return out if like_torch else out.astype(np.asarray(x).dtype if np.asarray(x).dtype.kind == "f" else np.float32)

    def _reconstruct_real(self, coeffs):
        like_torch = dv.is_torch(coeffs)
        eng = self.device_engine()
        if coeffs.shape[0] != eng.ncoef:
            raise ValueError("expected %d coefficients, got %d" % (eng.ncoef, coeffs.shape[0]))
        c, pix = dv.pack_real(coeffs, eng.ncoef, eng.ldc, eng.device)
        sig = eng.reconstruct(c)
        out = dv.unpack_real(sig, eng.N, pix, like_torch)
        return out if like_torch else out.astype(np.asarray(coeffs).dtype if np.asarray(coeffs).dtype.kind == "f" else np.float32)

    # ---- reference API --------------------------------------------------------------------------------------
    def decompose(self, x):
        self._check_level(x.shape[0])
        return self._decompose_real(x)

    def reconstruct(self, input_coeffs):
        return self._reconstruct_real(input_coeffs)

    def decompose_complex(self, x):
        self._check_level(x.shape[0])
        re = self._decompose_real(x.real)
        im = self._decompose_real(x.imag)
        return torch.complex(re, im) if dv.is_torch(x) else re + 1j * im

    def reconstruct_complex(self, input_coeffs):
        re = self._reconstruct_real(input_coeffs.real)
        im = self._reconstruct_real(input_coeffs.imag)
        return torch.complex(re, im) if dv.is_torch(input_coeffs) else re + 1j * im

    def _check_level(self, length):
        if self.wavelet_level is not None and self.wavelet_level > self._max_level_of(length):
            raise ValueError("You are trying to decompose into more levels than the maximum level expected")

    def _max_level_of(self, length):
        raise NotImplementedError
