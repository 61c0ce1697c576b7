from .reconstructors import CSROMERReconstructorWrapper, FaradayReconstructorWrapper

__all__ = ["CSROMERReconstructorWrapper", "FaradayReconstructorWrapper"]
