from .flagger import Flagger, median_absolute_deviation, moving_average
from .hampel_flagger import HampelFlagger, hampel
from .manual_flagger import ManualFlagger
from .mean_flagger import MeanFlagger, normal_flagging

__all__ = ["Flagger", "MeanFlagger", "HampelFlagger", "ManualFlagger", "hampel", "normal_flagging",
           "median_absolute_deviation", "moving_average"]
