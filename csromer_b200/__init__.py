"""csromer_b200 -- B200-native implementation of CS-ROMER's reconstruction hot path.

Same public classes as the reference (``Dataset``, ``Parameter``, ``NDFT1D``/``DFT1D``, ``NUFFT1D``, ``Chi2``,
``L1``, ``OFunction``, ``FISTA``, ``DiscreteWavelet``, ``UndecimatedWavelet``, the synthetic sources), with the
lambda^2 <-> phi transforms, the Chi2 gradient, the L1 / wavelet proximal steps and the FISTA loop running on
sm_100a CUDA kernels for every line of sight of a cube at once.  See DESIGN.md.
"""
from ._device import SparseRows
from .base import Dataset
from .cube import estimate_lipschitz, reconstruct_cube
from .faraday_sky import FaradaySky
from .dictionaries import DiscreteWavelet, UndecimatedWavelet, Wavelet
from .objectivefunction import L1, Chi2, Fi, OFunction
from .optimization import FISTA, Optimizer
from .reconstruction import Parameter
from .simulation import FaradaySource, FaradayThickSource, FaradayThinSource
from .io import DeviceCube, Reader, Writer, filter_cubes
from .transformers import (DFT1D, FT, NDFT1D, NUFFT1D, DatasetStack, Flagger, HampelFlagger, ManualFlagger, MeanFlagger,
                           PerPixelDFT1D)
from .utils import calculate_noise
from .wrappers import CSROMERReconstructorWrapper, FaradayReconstructorWrapper

__version__ = "0.1.0"
__all__ = [
    "Dataset", "Parameter", "FT", "NDFT1D", "DFT1D", "NUFFT1D", "PerPixelDFT1D", "DatasetStack", "Wavelet", "DiscreteWavelet", "UndecimatedWavelet",
    "Fi", "Chi2", "L1", "OFunction", "Optimizer", "FISTA", "FaradaySource", "FaradayThinSource",
    "FaradayThickSource", "reconstruct_cube", "estimate_lipschitz", "Flagger", "MeanFlagger", "HampelFlagger",
    "ManualFlagger", "FaradaySky", "SparseRows", "CSROMERReconstructorWrapper", "FaradayReconstructorWrapper", "Reader", "Writer", "filter_cubes", "DeviceCube", "calculate_noise",
]
