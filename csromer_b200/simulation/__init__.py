from .faradaysource import FaradaySource
from .thicksource import FaradayThickSource
from .thinsource import FaradayThinSource

__all__ = ["FaradaySource", "FaradayThinSource", "FaradayThickSource"]
