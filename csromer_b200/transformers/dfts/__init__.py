from .ft import FT
from .ndft import DFT1D, NDFT1D
from .nufft import NUFFT1D
from .perpixel import DatasetStack, PerPixelDFT1D

__all__ = ["FT", "NDFT1D", "DFT1D", "NUFFT1D", "PerPixelDFT1D", "DatasetStack"]
