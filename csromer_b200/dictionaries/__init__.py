from .discrete import DiscreteWavelet
from .undecimated import UndecimatedWavelet
from .wavelet import Wavelet

__all__ = ["Wavelet", "DiscreteWavelet", "UndecimatedWavelet"]
