"""``PerPixelDFT1D`` -- the lambda^2 <-> phi operator for a stack of lines of sight that do NOT share their
lambda^2 sampling (kernel K4).

The reference has no batch API: the README cube recipe builds one ``Dataset`` + ``DFT1D`` / ``NUFFT1D`` per pixel
(README.md:327-332), and its flaggers can *delete* channels of a pixel (transformers/flaggers/mean_flagger.py:36-40,
simulation/faradaysource.py:65-112), after which that pixel has its own ``lambda2``, ``m``, ``s``, ``k`` and
``l2_ref``.  This class takes exactly those per-pixel ``Dataset`` objects (a sequence) plus the shared
``Parameter`` (uniform phi grid) and evaluates the same operator as ``NDFT1D`` (transformers/dfts/ndft.py:17-43) /
``NUFFT1D`` (nufft.py:50-77) for all of them in one launch per transform, twiddles generated on the fly.

Arrays cross this API channel-first like everywhere else: ``(m_max, P)`` zero-padded past each line of sight's own
channel count, ``(n, P)`` in Faraday depth.  ``operator.dataset`` is a ``DatasetStack`` view that behaves like one
batched ``Dataset`` for ``Chi2`` / ``FISTA`` (assigning ``model_data`` fans out to the member datasets).
"""
from dataclasses import dataclass

import numpy as np
import torch

from ... import _device as dv
from .ft import FT


class DatasetStack:
    """A sequence of 1-D ``Dataset`` objects presented as one batch (channel axis first, zero padded)."""

    def __init__(self, datasets):
        self.members = list(datasets)
        if not self.members:
            raise ValueError("DatasetStack needs at least one dataset")
        for ds in self.members:
            if np.ndim(ds.data) != 1 or np.ndim(ds.w) != 1:
                raise ValueError("DatasetStack members must be single lines of sight")
        self.m_count = np.array([ds.m for ds in self.members], dtype=np.int32)
        self.m = int(self.m_count.max())
        self.P = len(self.members)

    def _stack(self, getter, fill=0.0, dtype=np.float64):
        out = np.full((self.m, self.P), fill, dtype=dtype)
        for p, ds in enumerate(self.members):
            out[:ds.m, p] = getter(ds)
        return out

    @property
    def data(self):
        return self._stack(lambda ds: ds.data, dtype=np.complex128)

    @property
    def w(self):
        return self._stack(lambda ds: ds.w)

    @property
    def s(self):
        return self._stack(lambda ds: ds.s, fill=1.0)

    @property
    def sigma(self):
        return self._stack(lambda ds: ds.sigma, fill=np.inf)

    @property
    def lambda2_minus_ref(self):
        return self._stack(lambda ds: np.asarray(ds.lambda2, dtype=np.float64) - float(ds.l2_ref))

    @property
    def k(self):
        return np.array([float(ds.k) for ds in self.members])

    @property
    def theo_noise(self):
        return np.array([np.nan if ds.theo_noise is None else float(ds.theo_noise) for ds in self.members])

    # model_data / residual: fan out to the members (each member's setter refreshes its own residual)
    @property
    def model_data(self):
        self._resolve_deferred()
        return self._stack(lambda ds: ds.model_data, dtype=np.complex128)

    @model_data.setter
    def model_data(self, val):
        val = np.asarray(val)
        for p, ds in enumerate(self.members):
            ds.model_data = val[:ds.m, p].astype(np.complex64)

    @property
    def residual(self):
        self._resolve_deferred()
        return self._stack(lambda ds: ds.residual, dtype=np.complex128)

    def defer_model(self, model_thunk, residual_thunk):
        """the fused loop's final model: users read it from the MEMBER datasets, so it is fanned out right away"""
        self.model_data = model_thunk()

    def _resolve_deferred(self):
        pass


@dataclass(init=True, repr=True)
class PerPixelDFT1D(FT):
    """``dataset``: a sequence of per-line-of-sight ``Dataset`` objects (or a ``DatasetStack``)."""

    def __post_init__(self):
        super().__post_init__()
        if self.dataset is not None and not isinstance(self.dataset, DatasetStack):
            self.dataset = DatasetStack(self.dataset)
        self._stack = None

    def configure(self):
        self.device_stack()

    def device_stack(self):
        if self.dataset is None or self.parameter is None or self.parameter.phi is None:
            raise ValueError("operator needs datasets and a parameter with a phi grid")
        if self._stack is None:
            self._stack = dv.DeviceStack(self.dataset.lambda2_minus_ref.T, self.dataset.m_count, self.parameter.phi)
        return self._stack

    device_plan = device_stack  # what FISTA's fused path asks for

    def _chan(self, values, st):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(values, dtype=np.float64).T), dtype=torch.float32).to(st.device)

    def _adjoint(self, b, row_scale_value):
        st = self.device_stack()
        bt, _ = dv.pack_planar(np.asarray(b).reshape(st.m, st.P), st.m, st.ld2m, st.device)
        chan = self._chan(self.dataset.w / self.dataset.s, st)
        row = dv.per_pixel_vector(row_scale_value, st.P, st.device, "row scale")
        return dv.unpack_complex(st.adjoint(bt, chan, row), st.n, (st.P,), like_torch=dv.is_torch(b))

    def _forward(self, x, normalized):
        st = self.device_stack()
        xt, pix = dv.pack_planar(x, st.n, st.ld2n, st.device)
        if xt.shape[0] == 1 and st.P > 1:
            xt = xt.expand(st.P, st.ld2n).contiguous()
        chan = row = None
        if normalized:
            w, s = self.dataset.w, self.dataset.s
            factor = np.where(w > 0.0, s / (np.where(w > 0.0, w, 1.0) * st.n), 0.0)
            chan = self._chan(factor, st)
            row = dv.per_pixel_vector(self.dataset.k, st.P, st.device, "k")
        return dv.unpack_complex(st.forward(xt, chan, row), st.m, (st.P,), like_torch=dv.is_torch(x))

    def forward(self, x):
        return self._forward(x, normalized=False)

    def forward_normalized(self, x):
        return self._forward(x, normalized=True)

    def backward(self, b):
        return self._adjoint(b, 1.0 / self.dataset.k)

    def RMTF(self, phi_x=0.0):
        ones = (np.arange(self.dataset.m)[:, None] < self.dataset.m_count[None, :]).astype(np.complex64)
        return self._adjoint(ones, 1.0 / self.dataset.k)
