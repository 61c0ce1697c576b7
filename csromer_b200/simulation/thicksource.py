"""``FaradayThickSource`` -- a top-hat (Burn slab) component (reference: simulation/thicksource.py:10-26)."""
import numpy as np

from .faradaysource import FaradaySource


class FaradayThickSource(FaradaySource):

    def __init__(self, phi_fg=None, phi_center=None, **kwargs):
        super().__init__(**kwargs)
        self.phi_fg = phi_fg
        self.phi_center = 0.0 if phi_center is None else phi_center

    def simulate(self):
        """s_nu (nu/nu_0)^alpha exp(2i lambda2 phi_center) sinc(phi_fg lambda2 / pi)"""
        centre = np.multiply.outer(self.lambda2, np.asarray(self.phi_center, dtype=np.float64))
        width = np.multiply.outer(self.lambda2, np.asarray(self.phi_fg, dtype=np.float64))
        amp = np.multiply.outer(self.s, np.asarray(self.s_nu, dtype=np.float64))
        self.data = amp * np.exp(2j * centre) * np.sinc(width / np.pi)
