from .csromer_reconstructor import CSROMERReconstructorWrapper, sigma_clipped_std
from .faraday_reconstructor import FaradayReconstructorWrapper

__all__ = ["CSROMERReconstructorWrapper", "FaradayReconstructorWrapper", "sigma_clipped_std"]
