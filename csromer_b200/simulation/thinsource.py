"""``FaradayThinSource`` -- a single Faraday-thin component (reference: simulation/thinsource.py:10-29)."""
import numpy as np

from .faradaysource import FaradaySource, per_pixel_columns


class FaradayThinSource(FaradaySource):

    def __init__(self, phi_gal=None, dchi=None, **kwargs):
        super().__init__(**kwargs)
        self.phi_gal = phi_gal
        self.dchi = 0.0 if dchi is None else dchi

    def simulate(self):
        """s_nu (nu/nu_0)^alpha [cos(2 phi lambda2) + i sin(2 phi lambda2 + dchi)]; ``phi_gal`` / ``s_nu`` may be
        per-pixel arrays (same pixel shape), giving data of shape (m, *pix)."""
        l2, s, (phi, flux) = per_pixel_columns(self.lambda2, self.s, self.phi_gal, self.s_nu)
        phase = 2.0 * phi * l2
        self.data = flux * s * (np.cos(phase) + 1j * np.sin(phase + self.dchi))
