"""``UndecimatedWavelet`` -- stationary wavelet transform dictionary (reference: dictionaries/undecimated.py:10-171).

Semantics kept: ``pywt.swt(trim_approx=True)`` analysis to ``level`` (default: as many levels as the length halves
exactly), coefficients flattened ``[cA_L, cD_L, ..., cD_1]`` (``pywt.ravel_coeffs`` order), optional signal in
front (``append_signal``), ``pywt.iswt`` synthesis, ``norm`` = PyWavelets' energy-preserving scaling (analysis
filters / sqrt(2), synthesis filters * sqrt(2)), and the reference's own "IUWT" bank.  Not kept: padding for a
length that is not a multiple of 2**level with ``mode`` other than zero padding.
"""
from dataclasses import dataclass

import numpy as np

from ..utils import next_power_2
from . import filters
from .wavelet import Wavelet


@dataclass(init=True, repr=True)
class UndecimatedWavelet(Wavelet):
    trim_approx: bool = None
    norm: bool = None

    _kind = 0

    def __post_init__(self):
        super().__post_init__()
        self.wavelet = filters.filter_bank(self.wavelet_name, self.filter_bank)
        self.trim_approx = True if self.trim_approx is None else self.trim_approx
        if not self.trim_approx:
            raise NotImplementedError("only trim_approx=True (the reference default) is implemented")
        self.norm = False if self.norm is None else self.norm
        if self.wavelet_level is not None:
            self.array_size = 2 ** self.wavelet_level
        self.pad_width = None

    @staticmethod
    def calculate_max_level(x):
        return filters.swt_max_level(len(x))

    def _max_level_of(self, length):
        return filters.swt_max_level(length)

    def _level_for(self, length):
        return filters.swt_max_level(length) if self.wavelet_level is None else self.wavelet_level

    def _norm_flag(self):
        return bool(self.norm)

    def decompose(self, x):
        length = x.shape[0]
        self._check_level(length)
        level = self._level_for(length)
        if length and length % (2 ** level) != 0:
            if self.mode is not None:
                raise NotImplementedError("padding mode %r is not implemented (only zero padding)" % (self.mode,))
            print("Your signal length is not multiple of 2**" + str(self.wavelet_level) + ". Padding array...")
            padded = next_power_2(length)
            self.pad_width = padded - length
            pad = [(0, self.pad_width)] + [(0, 0)] * (np.ndim(x) - 1)
            xp = np.pad(np.asarray(x), pad)
            coeffs = self._decompose_real(xp)
            if self.append_signal:  # the reference prepends the UNPADDED signal
                coeffs = np.concatenate([np.asarray(x), coeffs[padded:]], axis=0)
            self.n = length
            return coeffs
        return self._decompose_real(x)

    def reconstruct(self, input_coeffs):
        if self.pad_width is not None:
            length, padded = self.n, self.n + self.pad_width
            eng = self.prepare(padded)
            body = input_coeffs[length:] if self.append_signal else input_coeffs
            if self.append_signal:
                zeros = np.zeros((padded,) + tuple(np.shape(body)[1:]), dtype=np.asarray(body).dtype)
                full = np.concatenate([zeros, body], axis=0)
            else:
                full = body
            out = self._reconstruct_real(full)[:length]
            if self.append_signal:
                out = out + input_coeffs[:length]
            self.n = length
            self.pad_width = None
            return out
        return self._reconstruct_real(input_coeffs)
