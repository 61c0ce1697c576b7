"""``FT`` -- abstract lambda^2 <-> phi operator (reference: transformers/dfts/ft.py:23-51).

Concrete operators hold a REFERENCE to the dataset (whose ``model_data`` / ``residual`` the objective function
mutates) and a DEEP COPY of the parameter taken at construction, exactly like the reference, so later changes
of ``parameter.data`` do not touch the operator's phi grid.

The shared machinery lives here: both ``NDFT1D`` and ``NUFFT1D`` evaluate the same exact operator
E[i, j] = exp(2i (lambda2_i - l2_ref) phi_j) on the GPU through ``DevicePlan`` (one dense tensor-core GEMM over
every line of sight, per-channel / per-pixel diagonals fused around it).
"""
import copy
from abc import ABCMeta, abstractmethod
from dataclasses import dataclass

import weakref

import numpy as np
import torch

from ... import _device as dv
from ...base.dataset import Dataset
from ...reconstruction.parameter import Parameter


@dataclass(init=True, repr=True)
class FT(metaclass=ABCMeta):
    dataset: Dataset = None
    parameter: Parameter = None

    def __post_init__(self):
        if self.parameter is not None:
            self.parameter = copy.deepcopy(self.parameter)

    @abstractmethod
    def configure(self):
        pass

    @abstractmethod
    def forward(self, x):
        pass

    @abstractmethod
    def forward_normalized(self, x):
        pass

    @abstractmethod
    def backward(self, b):
        pass

    @abstractmethod
    def RMTF(self):
        pass

    # ---- device helpers shared by the concrete operators ------------------------------------------------
    def device_plan(self):
        """the cached twiddle tables for this (lambda2, l2_ref, phi) sampling"""
        if self.dataset is None or self.parameter is None or self.parameter.phi is None:
            raise ValueError("operator needs a dataset and a parameter with a phi grid")
        return dv.DevicePlan.get(self.dataset.lambda2, self.dataset.l2_ref, self.parameter.phi)

    def _channel_tensor(self, values, plan):
        """per-channel ``(m,)`` or per-pixel ``(m, *pix)`` host array -> fp32 ``[m]`` / ``[P, m]`` on the device"""
        v = np.asarray(values, dtype=np.float64)
        if v.ndim == 1:
            return dv.small_to_device(v, torch.float32, plan.device)
        return dv.small_to_device(v.reshape(plan.m, -1).T, torch.float32, plan.device)

    def _row_tensor(self, values, P, plan, name="k"):
        return dv.per_pixel_vector(values, P, plan.device, name)

    def _weights(self):
        ds = self.dataset
        w = np.asarray(ds.w, dtype=np.float64)
        s = np.asarray(ds.s, dtype=np.float64)
        return w, s

    def _adjoint(self, b, row_scale_value, remember=False):
        """row_scale * E^H ((w/s) b) for ``b`` complex ``(m, *pix)``.  ``remember`` (backward of the dataset's own torch
        data): the planar device result is kept next to the returned tensor, so that a FISTA run started from exactly this
        dirty spectrum (README.md:183-190) neither transposes it back nor computes it a second time (``dirty_of``)."""
        plan = self.device_plan()
        if b is self.dataset.data and not isinstance(b, np.ndarray):    # one upload per dataset (Dataset.device_data)
            bt, pix = self.dataset.device_data(plan.device)
        else:
            bt, pix = dv.pack_planar(b, plan.m, plan.ld2m, plan.device)
        w, s = self._weights()
        s_col = s if w.ndim == 1 else s.reshape((-1,) + (1,) * (w.ndim - 1))
        chan = self._channel_tensor(w / s_col, plan)
        row = self._row_tensor(row_scale_value, bt.shape[0], plan)
        out = plan.adjoint(bt, chan, row)
        res = dv.unpack_complex(out, plan.n, pix, like_torch=dv.is_torch(b))
        self._dirty_cache = None
        if remember and b is self.dataset.data and dv.is_torch(res):
            ds = self.dataset
            self._dirty_cache = dict(res=weakref.ref(res), version=res._version, planar=out, data=weakref.ref(b), data_version=b._version,
                                     w=ds.w, s=ds.s, k=ds.k, plan=plan)
        return res

    def dirty_of(self, spectrum):
        """the planar device copy of ``spectrum`` if that tensor is the untouched result of ``backward(dataset.data)`` for the
        dataset as it is now, else None"""
        c = getattr(self, "_dirty_cache", None)
        if c is None or spectrum is None or c["res"]() is not spectrum or spectrum._version != c["version"]:
            return None
        ds = self.dataset
        data = c["data"]()
        if data is None or data is not ds.data or data._version != c["data_version"]:
            return None
        if ds.w is not c["w"] or ds.s is not c["s"] or ds.k is not c["k"] or self.device_plan() is not c["plan"]:
            return None
        return c["planar"]

    def _forward(self, x, normalized):
        plan = self.device_plan()
        xt, pix = dv.pack_planar(x, plan.n, plan.ld2n, plan.device)
        chan = row = None
        if normalized:
            w, s = self._weights()
            s_col = s if w.ndim == 1 else s.reshape((-1,) + (1,) * (w.ndim - 1))
            with np.errstate(divide="ignore", invalid="ignore"):
                factor = np.where(w > 0.0, s_col / (np.where(w > 0.0, w, 1.0) * plan.n), 0.0)
            chan = self._channel_tensor(factor, plan)
            row = self._row_tensor(self.dataset.k, xt.shape[0], plan)
        out = plan.forward(xt, chan, row)
        return dv.unpack_complex(out, plan.m, pix, like_torch=dv.is_torch(x))
