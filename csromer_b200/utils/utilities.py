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


def make_mask(I=None, sigma=0.0):
    """(indices where ``I >= sigma``, indices where ``I < sigma``) as ``np.where`` tuples (utils/utilities.py:35-38)"""
    I = np.asarray(I)
    keep = I >= sigma
    return np.where(keep), np.where(I < sigma)


def make_mask_faraday(I=None, P=None, cube_Q=None, cube_U=None, spectral_idx=None, sigma_I=0.0, sigma_P=0.0):
    """Pixels worth reconstructing: total intensity and polarised intensity above their thresholds and not a single NaN
    along the frequency axis of Q or U (nor in the spectral-index map when one is given).  Returns ``(kept, masked)`` as
    ``np.where`` tuples like the reference (utils/utilities.py:41-65); the cube recipe loops over ``kept``
    (README.md:298-307)."""
    keep = (np.asarray(I) >= sigma_I) & (np.asarray(P) >= sigma_P)
    for cube in (cube_Q, cube_U):
        if cube is not None:
            keep &= ~np.isnan(np.asarray(cube)).any(axis=0)
    if spectral_idx is not None:
        keep &= ~np.isnan(np.asarray(spectral_idx))
    return np.where(keep), np.where(~keep)


def calculate_noise(image=None, x0=None, xn=None, y0=None, yn=None, nsigma=3, cenfunc="mean", stdfunc="mad_std",
                    use_sigma_clipped_stats=False):
    """Noise of an image ``(Y, X)`` or of every channel of a cube ``(m, Y, X)`` over the window ``[y0:yn, x0:xn]``:
    ``numpy.nanstd`` (reference: utils/utilities.py:68-110).  A CUDA tensor (or a ``DeviceCube`` plane stack) is reduced
    on the GPU by ``csr_channel_stats``; host arrays take the NumPy path the reference takes.  The sigma-clipped variant
    needs astropy (``sigma_clipped_stats``), which is not available here."""
    if use_sigma_clipped_stats:
        raise NotImplementedError("sigma-clipped statistics need astropy; use use_sigma_clipped_stats=False")
    nd = image.dim() if (torch is not None and isinstance(image, torch.Tensor)) else np.ndim(image)
    Y, X = (image.shape[1], image.shape[2]) if nd > 2 else (image.shape[0], image.shape[1])
    x0, y0 = 0 if x0 is None else x0, 0 if y0 is None else y0
    xn, yn = X if xn is None else xn, Y if yn is None else yn
    if torch is not None and isinstance(image, torch.Tensor) and image.is_cuda:
        from .. import _device as dv
        from .. import _lib
        cube = image.reshape((-1, Y, X)).to(torch.float32).contiguous()
        stats = torch.empty((cube.shape[0], 2), dtype=torch.float32, device=cube.device)
        _lib.check(_lib.load().csr_channel_stats(dv.tptr(cube), cube.shape[0], Y, X, int(x0), int(xn), int(y0), int(yn),
                                                 dv.tptr(stats), dv.stream_ptr()), "csr_channel_stats")
        return stats[:, 1] if nd > 2 else stats[0, 1]
    image = np.asarray(image)
    if nd > 2:
        return np.nanstd(image[:, y0:yn, x0:xn], axis=(1, 2))
    return np.nanstd(image[y0:yn, x0:xn])
