from .utilities import calculate_noise, complex_to_real, make_mask, make_mask_faraday, next_power_2, real_to_complex

__all__ = ["calculate_noise", "complex_to_real", "make_mask", "make_mask_faraday", "next_power_2", "real_to_complex"]
