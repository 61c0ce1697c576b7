from .faraday_sky import FaradaySky

__all__ = ["FaradaySky"]
