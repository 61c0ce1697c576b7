"""``DiscreteWavelet`` -- decimated wavelet dictionary (reference: dictionaries/discrete.py:10-130).

Semantics kept: ``pywt.wavedec`` / ``pywt.waverec`` with ``mode="periodization"`` (the mode the reference README
uses, README.md:227) or any of PyWavelets' signal-extension modes (``zero``, ``constant``, ``symmetric``, ``reflect``,
``periodic``, ``smooth``, ``antisymmetric``, ``antireflect``: band length ``(N + F - 1) // 2`` per level), levels =
``dwt_max_level`` by default, coefficients flattened ``[cA_L, cD_L, ..., cD_1]`` with odd-length levels handled like
PyWavelets (periodization: last sample repeated on analysis, extra sample dropped on synthesis; the running
approximation is cut to the next detail band's length in every mode), optional ``append_signal``.  ``mode=None`` is an
error here as it is in PyWavelets.  One difference: ``reconstruct`` returns the recorded signal length ``n`` -- for an
ODD-length signal in an extension mode ``pywt.waverec`` returns ``n + 1`` samples (the planar vectors of this package
have even length, so the hot path never sees it).
"""
from dataclasses import dataclass

from . import filters
from .wavelet import Wavelet


@dataclass(init=True, repr=True)
class DiscreteWavelet(Wavelet):

    _kind = 1
    # ``kind`` argument of csr_wavelet_create (include/csromer_b200.h)
    MODES = {"periodization": 1, "zero": 2, "constant": 3, "symmetric": 4, "reflect": 5, "periodic": 6, "smooth": 7,
             "antisymmetric": 8, "antireflect": 9}

    def __post_init__(self):
        super().__post_init__()
        if self.wavelet_name == "IUWT":
            raise NotImplementedError("The wavelet has not been implemented by pywavelets")
        self.wavelet = filters.filter_bank(self.wavelet_name, self.filter_bank)
        if not isinstance(self.mode, str):
            raise TypeError("DiscreteWavelet: mode must be the name of a PyWavelets signal-extension mode (got %r)" % (self.mode,))
        if self.mode not in self.MODES:
            raise ValueError("Unknown mode name %r; one of %s" % (self.mode, ", ".join(sorted(self.MODES))))
        self._kind = self.MODES[self.mode]

    def calculate_ncoeffs(self, x):
        """pywt.dwt_coeff_len(len(x), filter length, mode) (reference: discrete.py:20-22)"""
        if self.mode == "periodization":
            return (len(x) + 1) // 2
        return (len(x) + len(self.wavelet[0]) - 1) // 2

    def calculate_max_level(self, x):
        return filters.dwt_max_level(len(x), len(self.wavelet[0]))

    def _max_level_of(self, length):
        return filters.dwt_max_level(length, len(self.wavelet[0]))

    def _level_for(self, length):
        return self._max_level_of(length) if self.wavelet_level is None else self.wavelet_level
