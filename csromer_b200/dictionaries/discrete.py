"""``DiscreteWavelet`` -- decimated wavelet dictionary (reference: dictionaries/discrete.py:10-130).

Semantics kept: ``pywt.wavedec`` / ``pywt.waverec`` with ``mode="periodization"`` (the mode the reference README
uses, README.md:227), levels = ``dwt_max_level`` by default, coefficients flattened ``[cA_L, cD_L, ..., cD_1]``
with odd-length levels handled like PyWavelets (last sample repeated on analysis, extra sample dropped on
synthesis), optional ``append_signal``.  Other signal-extension modes are not implemented; ``mode=None`` is an
error here as it is in PyWavelets.
"""
from dataclasses import dataclass

from . import filters
from .wavelet import Wavelet


@dataclass(init=True, repr=True)
class DiscreteWavelet(Wavelet):

    _kind = 1

    def __post_init__(self):
        super().__post_init__()
        if self.wavelet_name == "IUWT":
            raise NotImplementedError("The wavelet has not been implemented by pywavelets")
        self.wavelet = filters.filter_bank(self.wavelet_name, self.filter_bank)
        if self.mode != "periodization":
            raise NotImplementedError("DiscreteWavelet: only mode='periodization' is implemented (got %r)" % (self.mode,))

    def calculate_ncoeffs(self, x):
        return (len(x) + 1) // 2  # pywt.dwt_coeff_len(..., mode="periodization")

    def calculate_max_level(self, x):
        return filters.dwt_max_level(len(x), len(self.wavelet[0]))

    def _max_level_of(self, length):
        return filters.dwt_max_level(length, len(self.wavelet[0]))

    def _level_for(self, length):
        return self._max_level_of(length) if self.wavelet_level is None else self.wavelet_level
