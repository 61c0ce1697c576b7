"""Small host helpers kept from the reference API (utils/utilities.py:12-32 in the reference)."""
import numpy as np

try:  # torch tensors are accepted wherever arrays are
    import torch
except ImportError:  # pragma: no cover
    torch = None


def next_power_2(n):
    """Smallest power of two >= n (n itself when it already is one); 1 for n == 0."""
    n = int(n)
    if n > 0 and n & (n - 1) == 0:
        return n
    return 1 << n.bit_length()


def real_to_complex(z):
    """planar real ``[Re | Im]`` of length 2n along axis 0 -> complex of length n"""
    half = z.shape[0] // 2
    if torch is not None and isinstance(z, torch.Tensor):
        return torch.complex(z[:half], z[half:])
    return z[:half] + 1j * z[half:]


def complex_to_real(z):
    """complex of length n along axis 0 -> planar real ``[Re | Im]`` of length 2n"""
    if torch is not None and isinstance(z, torch.Tensor):
        return torch.cat([z.real, z.imag], dim=0)
    return np.concatenate([z.real, z.imag], axis=0)
