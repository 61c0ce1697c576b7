"""``FaradaySource`` -- synthetic lambda^2-space sources used as test and benchmark input
(reference: simulation/faradaysource.py:13-143).  Host-side; not on the hot path.
"""
import copy

import numpy as np

from ..base.dataset import Dataset


def per_pixel_columns(lambda2, s, *params):
    """Broadcast helper for batched sources: returns lambda2 and s as columns ``(m, 1, ...)`` and the parameters as
    float64 arrays of one common pixel shape (scalars stay 0-d, so a single line of sight stays 1-D)."""
    arrays = np.broadcast_arrays(*[np.asarray(p, dtype=np.float64) for p in params])
    extra = (1,) * arrays[0].ndim
    return lambda2.reshape((-1,) + extra), np.asarray(s).reshape((-1,) + extra), arrays


class FaradaySource(Dataset):

    def __init__(self, s_nu=None, **kwargs):
        super().__init__(**kwargs)
        self.s_nu = s_nu
        self.sigma = np.ones_like(self.lambda2)

    def simulate(self):
        raise NotImplementedError

    def _combined(self, other, duplicate):
        if not isinstance(other, FaradaySource):
            return NotImplemented
        if self.data is None or other.data is None or self.s_nu is None or other.s_nu is None \
                or not np.array_equal(self.nu, other.nu):
            raise TypeError("Data attribute in sources cannot be NoneType")
        total = duplicate(self)
        total.data = self.data + other.data
        flux = self.s_nu + other.s_nu
        # flux-weighted spectral index; the result keeps the class of the LEFT operand (reference behaviour)
        total.spectral_idx = (self.s_nu * self.spectral_idx + other.s_nu * other.spectral_idx) / flux
        return total

    def __add__(self, other):
        return self._combined(other, copy.deepcopy)

    def __iadd__(self, other):
        return self._combined(other, copy.copy)

    def add_external_faraday_depolarization(self, sigma_rm=None):
        sigma_rm = 0.0 if sigma_rm is None else sigma_rm
        self.data = self.data * np.exp(-2.0 * sigma_rm ** 2 * self.lambda2 ** 2)

    def remove_channels(self, remove_frac=None, random_state=None, chunksize=None):
        """Keep a random ~(1 - remove_frac) subset of the channels, removed in RFI-like contiguous chunks.
        Channels are DELETED (m changes), as in the reference (faradaysource.py:65-112)."""
        if not remove_frac:
            return
        rng = np.random if random_state is None else random_state
        keep_frac = 1.0 - remove_frac
        if chunksize is None:
            chunksize = self.m // (10 + (10 * (1.0 - keep_frac)))
        kept = np.zeros(self.m, dtype=bool)
        while kept.mean() <= keep_frac:
            pos = rng.randint(0, self.m)
            width = rng.uniform(0, chunksize)
            lo = max(int(pos - 0.5 * width), 0)
            hi = min(int(pos + 0.5 * width), self.m - 1)
            kept[lo:hi] = True
        chans = np.flatnonzero(kept)
        while len(chans) / float(self.m) > keep_frac:
            chans = np.delete(chans, rng.randint(0, len(chans)))
        data = self.data
        self.lambda2 = self.lambda2[chans]
        if data is not None:
            self.data = data[chans]

    def apply_noise(self, noise=None, random_state=None):
        if noise is None:
            return
        if isinstance(noise, complex):
            sig_q, sig_u = np.float32(noise.real), np.float32(noise.imag)
        elif isinstance(noise, float):
            sig_q = sig_u = np.float32(noise)
        else:
            raise TypeError("Noise must be either a float or complex number")
        mean_sigma = (sig_q + sig_u) / 2.0
        if not mean_sigma > 0.0:
            self.sigma = np.ones_like(self.lambda2)
            return
        self.sigma = np.ones_like(self.lambda2) * mean_sigma
        rng = np.random if random_state is None else random_state
        shape = np.shape(self.data)
        q = rng.normal(loc=0.0, scale=sig_q, size=shape)
        u = rng.normal(loc=0.0, scale=sig_u, size=shape)
        self.data = self.data + (q + 1j * u)
