from .utilities import calculate_noise, complex_to_real, next_power_2, real_to_complex

__all__ = ["calculate_noise", "complex_to_real", "next_power_2", "real_to_complex"]
