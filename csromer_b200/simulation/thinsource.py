"""``FaradayThinSource`` -- a single Faraday-thin component (reference: simulation/thinsource.py:10-29)."""
import numpy as np

from .faradaysource import FaradaySource


class FaradayThinSource(FaradaySource):

    def __init__(self, phi_gal=None, dchi=None, **kwargs):
        super().__init__(**kwargs)
        self.phi_gal = phi_gal
        self.dchi = 0.0 if dchi is None else dchi

    def simulate(self):
        """s_nu (nu/nu_0)^alpha [cos(2 phi lambda2) + i sin(2 phi lambda2 + dchi)]; ``phi_gal`` / ``s_nu`` may be
        per-pixel arrays, giving data of shape (m, *pix)."""
        phase = 2.0 * np.multiply.outer(self.lambda2, np.asarray(self.phi_gal, dtype=np.float64))
        amp = np.multiply.outer(self.s, np.asarray(self.s_nu, dtype=np.float64))
        self.data = amp * (np.cos(phase) + 1j * np.sin(phase + self.dchi))
