from .chi2 import Chi2
from .l1 import L1

__all__ = ["Chi2", "L1"]
