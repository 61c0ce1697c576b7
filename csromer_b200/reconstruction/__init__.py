from .parameter import Parameter

__all__ = ["Parameter"]
