"""``FaradayThickSource`` -- a top-hat (Burn slab) component (reference: simulation/thicksource.py:10-26)."""
import numpy as np

from .faradaysource import FaradaySource, per_pixel_columns


class FaradayThickSource(FaradaySource):

    def __init__(self, phi_fg=None, phi_center=None, **kwargs):
        super().__init__(**kwargs)
        self.phi_fg = phi_fg
        self.phi_center = 0.0 if phi_center is None else phi_center

    def simulate(self):
        """s_nu (nu/nu_0)^alpha exp(2i lambda2 phi_center) sinc(phi_fg lambda2 / pi); parameters may be per-pixel"""
        l2, s, (width, centre, flux) = per_pixel_columns(self.lambda2, self.s, self.phi_fg, self.phi_center, self.s_nu)
        self.data = flux * s * np.exp(2j * l2 * centre) * np.sinc(width * l2 / np.pi)
