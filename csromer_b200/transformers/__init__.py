from .dfts import DFT1D, FT, NDFT1D, NUFFT1D, DatasetStack, PerPixelDFT1D
from .flaggers import Flagger, HampelFlagger, ManualFlagger, MeanFlagger

__all__ = ["FT", "NDFT1D", "DFT1D", "NUFFT1D", "PerPixelDFT1D", "DatasetStack", "Flagger", "MeanFlagger", "HampelFlagger",
           "ManualFlagger"]
