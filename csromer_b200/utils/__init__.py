from .utilities import complex_to_real, next_power_2, real_to_complex

__all__ = ["complex_to_real", "next_power_2", "real_to_complex"]
