"""``Dataset`` -- the lambda^2-space container the operators and objective functions read and write.

Host-side mirror of the reference's ``base/dataset.py:84-381`` (same constructor, attribute names and setter
side effects), extended to a batch of lines of sight: every per-channel array keeps the channel axis FIRST,
so ``data`` may be ``(m,)``, ``(m, P)`` or ``(m, Y, X)`` and ``w`` / ``sigma`` either ``(m,)`` (shared) or
shaped like ``data`` (per-pixel weights and flags).  ``k`` and ``theo_noise`` are then per-pixel arrays.

Attribute semantics kept from the reference:
  * ``lambda2`` is stored ascending and ``nu`` is re-derived from it; ``data`` / ``sigma`` stay in caller order;
  * assigning ``lambda2`` resets ``w`` to ones; assigning ``sigma`` sets ``w = 1/sigma^2``; assigning ``w``
    refreshes ``sigma``, ``k = sum(w/s)`` and ``theo_noise`` (None when every weight is exactly 1);
  * writing into ``w`` in place (``ds.w[idx] = 0``, what the reference's flaggers do) bypasses the setter, so
    ``k`` stays as it was -- the operators always take ``k`` as held by the object;
  * assigning ``model_data`` recomputes ``residual = data - model_data``.
The residual diagnostics of the reference (autocorrelation, Ljung-Box, histograms) are out of scope.
"""
import numpy as np

SPEED_OF_LIGHT = 299792458.0


def _channel_column(v, like):
    """reshape a per-channel vector so it broadcasts against ``like`` (channel axis first)"""
    v = np.asarray(v)
    if v.ndim == 1 and getattr(like, "ndim", 1) > 1:
        return v.reshape((-1,) + (1,) * (like.ndim - 1))
    return v


class Dataset:

    def __init__(self, nu=None, lambda2=None, data=None, l2_ref=None, w=None, sigma=None, spectral_idx=None,
                 gridded=None):
        self._lambda2 = self._nu = self._nu_0 = None
        self._s = self._w = self._sigma = None
        self._data = self._model_data = None
        self._residual = None
        self._deferred = None  # (model thunk, residual thunk) left by the fused GPU loop, resolved on first read
        self._spectral_idx = 0.0
        self._m = None
        self.k = None
        self.theo_noise = None
        self.delta_l2_min = self.delta_l2_max = self.delta_l2_mean = 0.0
        self.l2_ref = 0.0 if l2_ref is None else l2_ref
        self.gridded = bool(gridded) if gridded is not None else False

        if lambda2 is not None:
            self.lambda2 = lambda2
        if nu is not None:
            self.nu = nu
        self.spectral_idx = spectral_idx
        if w is not None:
            self.w = w
        elif sigma is not None:
            self.sigma = sigma
        elif self._m is not None:
            self.sigma = np.ones(self._m)
        self.data = data
        if isinstance(self._data, np.ndarray) and self._model_data is None:
            self._model_data = np.zeros_like(self._data)

    # ---- sampling ------------------------------------------------------------------------------------
    @property
    def m(self):
        return self._m

    @m.setter
    def m(self, val):
        self._m = val

    @property
    def nu_0(self):
        return self._nu_0

    @nu_0.setter
    def nu_0(self, val):
        self._nu_0 = val

    @property
    def nu(self):
        return self._nu

    @nu.setter
    def nu(self, val):
        if val is None:
            self._nu = None
            return
        val = np.asarray(val, dtype=np.float64)
        # lambda2 ascending <=> frequency descending
        self.lambda2 = np.ascontiguousarray(((SPEED_OF_LIGHT / val) ** 2)[::-1])

    @property
    def lambda2(self):
        return self._lambda2

    @lambda2.setter
    def lambda2(self, val):
        if val is None:
            self._lambda2 = None
            return
        val = np.asarray(val, dtype=np.float64)
        if val.size > 1 and np.all(np.diff(val) < 0):
            val = np.ascontiguousarray(val[::-1])
        self._lambda2 = val
        self._m = len(val)
        self._nu = SPEED_OF_LIGHT / np.sqrt(val)
        self._nu_0 = 0.5 * (self._nu.min() + self._nu.max())
        self._s = (self._nu / self._nu_0) ** self._spectral_idx
        self.w = np.ones(self._m)
        self.calculate_l2_cellsize()

    @property
    def spectral_idx(self):
        return self._spectral_idx

    @spectral_idx.setter
    def spectral_idx(self, val):
        self._spectral_idx = 0.0 if val is None else val
        if self._lambda2 is not None and self._nu_0 is not None:
            self.s = (self._nu / self._nu_0) ** self._spectral_idx

    @property
    def s(self):
        return self._s

    @s.setter
    def s(self, val):
        self._s = val
        if val is not None and self._w is not None:
            self.k = self._weight_sum(self._w / _channel_column(val, self._w))

    # ---- weights ---------------------------------------------------------------------------------------
    @staticmethod
    def _weight_sum(v):
        total = np.sum(v, axis=0)
        return float(total) if np.ndim(total) == 0 else total

    @property
    def w(self):
        return self._w

    @w.setter
    def w(self, val):
        if val is None:
            self._w = None
            self.theo_noise = None
            return
        val = np.asarray(val, dtype=np.float64)
        self._w = val
        sig = np.zeros_like(val)
        nz = val != 0
        sig[nz] = 1.0 / np.sqrt(val[nz])
        self._sigma = sig
        if self._s is not None:
            self.k = self._weight_sum(val / _channel_column(self._s, val))
        else:
            self.k = self._weight_sum(val)
        if self.l2_ref is None:
            self.l2_ref = self.calculate_l2ref()
        self.theo_noise = self.calculate_theo_noise()

    @property
    def sigma(self):
        return self._sigma

    @sigma.setter
    def sigma(self, val):
        if val is not None:
            val = np.asarray(val, dtype=np.float64)
            self.w = 1.0 / (val ** 2)
        self._sigma = val

    # ---- data ------------------------------------------------------------------------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, val):
        if val is None:
            self._data = None
            return
        if self._m is None or len(val) != self._m:
            self._m = len(val)
        self._data = val
        if self._model_data is None:
            self._model_data = np.zeros_like(np.asarray(val)) if isinstance(val, np.ndarray) else None

    def __getstate__(self):  # (copies and pickles do not carry the device-side copy of the data)
        state = self.__dict__.copy()
        state.pop("_packed", None)
        return state

    # ---- device copy of the data (GPU path) -----------------------------------------------------------------
    def _data_key(self):
        d = self._data
        return (id(d), getattr(d, "_version", None), tuple(np.shape(d)))

    def device_data(self, device=None):
        """``data`` as the fp32 planar ``[P, ld2m]`` device matrix the kernels read, plus the pixel shape.  The copy is kept
        on this object, so the dirty spectrum and the FISTA run of one dataset share one host-to-device transfer; it is
        refreshed when ``data`` is re-assigned or (torch tensors) modified in place."""
        from .. import _device as dv
        held = getattr(self, "_packed", None)
        if held is not None and held[0] == self._data_key() and (device is None or held[1].device == dv.torch.device(device)):
            if held[3] is not None:
                dv.torch.cuda.current_stream().wait_event(held[3])     # uploaded by prefetch() on its own stream
            return held[1], held[2]
        planar, pix = dv.pack_planar(self._data, self._m, (2 * self._m + 7) // 8 * 8, device)
        if isinstance(self._data, np.ndarray) is False:                # NumPy arrays can change unnoticed: not cached
            self._packed = (self._data_key(), planar, pix, None)
        return planar, pix

    def prefetch(self, device=None):
        """Start the host-to-device transfer of ``data`` (pinned host tensor) on a side stream, so that it overlaps the
        reconstruction of the previous slab; ``device_data`` then waits for it instead of copying."""
        from .. import _device as dv
        torch = dv.torch
        stream = dv.side_stream(device)
        with torch.cuda.stream(stream):
            planar, pix = dv.pack_planar(self._data, self._m, (2 * self._m + 7) // 8 * 8, device)
            done = torch.cuda.Event()
            done.record(stream)
        self._packed = (self._data_key(), planar, pix, done)
        return self

    def defer_model(self, model_thunk, residual_thunk):
        """Used by the fused FISTA loop: ``model_data`` / ``residual`` of the final evaluation stay on the GPU until
        somebody reads them (for a cube they are large and often never read); the thunks convert to the type and
        device of ``data`` on first access.  Same observable values as assigning ``model_data`` directly."""
        self._deferred = (model_thunk, residual_thunk)
        self._model_data = self._residual = None

    def _resolve_deferred(self):
        if self._deferred is not None:
            model_thunk, residual_thunk = self._deferred
            self._deferred = None
            self._model_data = model_thunk()
            self._residual = residual_thunk()

    @property
    def residual(self):
        self._resolve_deferred()
        return self._residual

    @residual.setter
    def residual(self, val):
        self._deferred = None
        self._residual = val

    @property
    def model_data(self):
        self._resolve_deferred()
        return self._model_data

    @model_data.setter
    def model_data(self, val):
        self._deferred = None
        if val is None:
            self._model_data = None
            return
        if len(val) != self._m:
            raise SystemExit("Data must have same size as lambda2")  # reference: sys.exit (dataset.py:314)
        self._model_data = val
        if self._data is not None:
            self.calculate_residuals()

    @property
    def n_pixels(self):
        if self._data is None or np.ndim(self._data) < 2:
            return 1
        return int(np.prod(np.shape(self._data)[1:]))

    # ---- derived quantities ----------------------------------------------------------------------------
    def calculate_residuals(self):
        self.residual = self._data - self._model_data

    def calculate_l2ref(self):
        if self._lambda2 is None:
            return None
        l2 = _channel_column(self._lambda2, self._w)
        ref = np.sum(self._w * l2, axis=0) / np.sum(self._w, axis=0)
        return float(ref) if np.ndim(ref) == 0 else ref

    def calculate_l2_cellsize(self):
        if self._w is None or self._lambda2 is None:
            return
        good = self._w > 0.0 if self._w.ndim == 1 else np.any(self._w.reshape(self._m, -1) > 0.0, axis=1)
        spacing = np.abs(np.diff(self._lambda2[good]))
        if spacing.size:
            self.delta_l2_min = spacing.min()
            self.delta_l2_max = spacing.max()
            self.delta_l2_mean = spacing.mean()

    def calculate_theo_noise(self):
        if not isinstance(self._w, np.ndarray):
            return None
        if np.all(self._w == 1.0):
            return None
        tn = 1.0 / np.sqrt(np.sum(self._w, axis=0))
        return float(tn) if np.ndim(tn) == 0 else tn

    def calculate_amplitude(self, column="data"):
        if not hasattr(self, column):
            raise ValueError("Column does not exist")
        values = getattr(self, column)
        if not np.iscomplexobj(values):
            raise TypeError("Data is not complex")
        return np.abs(values)

    def calculate_polangle(self, column="data"):
        if not hasattr(self, column):
            raise ValueError("Column does not exist")
        values = getattr(self, column)
        if not np.iscomplexobj(values):
            raise TypeError("Data is not complex")
        # the reference always reads self.data here (dataset.py:339), whatever the column
        return 0.5 * np.arctan2(self._data.imag, self._data.real)

    def subtract_galacticrm(self, phi_gal):
        """d_i <- d_i exp(-2i phi_gal lambda2_i); ``phi_gal`` scalar or one value per pixel."""
        phase = np.multiply.outer(self._lambda2, np.asarray(phi_gal, dtype=np.float64))
        self.data = self._data * np.exp(-2j * phase.reshape((self._m,) + np.shape(self._data)[1:]))
