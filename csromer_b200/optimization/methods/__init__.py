from .fista import FISTA

__all__ = ["FISTA"]
