from .fi import Fi
from .ofunction import OFunction
from .priors import L1, Chi2

__all__ = ["Fi", "OFunction", "Chi2", "L1"]
