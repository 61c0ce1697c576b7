from .methods import FISTA
from .optimizer import Optimizer

__all__ = ["Optimizer", "FISTA"]
