"""Device plumbing between the reference-shaped Python objects and the C ABI.

PyTorch is used for device memory, streams and host<->device copies only; every arithmetic kernel on the
hot path lives in ``libcsromer_b200.so``.  Layout contract: user-facing arrays are
frequency / phi FIRST (``(m,)``, ``(m, P)``, ``(m, Y, X)``) like the reference's ``cube[:, i, j]``; on the device
everything is pixel-major planar fp32: ``[P, ld2m]`` and ``[P, ld2n]`` with ``[Re | Im]`` packing.
"""
import ctypes
import hashlib

import numpy as np
import torch

from . import _lib
from ._lib import CsrError


def require_cuda():
    if not torch.cuda.is_available():
        raise CsrError("csromer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def is_torch(a):
    return isinstance(a, torch.Tensor)


_side_streams = {}


def side_stream(device=None):
    """one extra stream per device for transfers that overlap the compute stream (Dataset.prefetch)"""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev not in _side_streams:
        _side_streams[dev] = torch.cuda.Stream(device=dev)
    return _side_streams[dev]


def pixel_shape(arr, lead):
    """(pixel shape tuple, P) of an array whose axis 0 has length ``lead``."""
    if arr.shape[0] != lead:
        raise ValueError("leading axis has length %d, expected %d" % (arr.shape[0], lead))
    pix = tuple(arr.shape[1:])
    return pix, int(np.prod(pix)) if pix else 1


def to_device(a, dtype=None, device=None):
    """numpy / torch -> torch tensor on the CUDA device (pinned staging for host arrays)."""
    require_cuda()
    if is_torch(a):
        t = a.to(device=device or "cuda", non_blocking=True)
    else:
        arr = np.ascontiguousarray(a)
        h = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
        t = h.pin_memory().to(device or "cuda", non_blocking=True)
    return t if dtype is None else t.to(dtype)


def _offset_ptr(t, byte_offset):
    return ctypes.c_void_p(t.data_ptr() + int(byte_offset))


def pack_planar(a, length, ld, device=None):
    """complex ``(length, *pix)`` or real ``(2*length, *pix)`` -> fp32 ``[P, ld]`` planar on the device.  The transpose
    is ``csr_pack_planar`` (32 x 32 shared-memory tiles), torch only moves the bytes to the device."""
    t = to_device(a, device=device)
    lib = _lib.load()
    if t.is_complex():
        pix, P = pixel_shape(t, length)
        t = t.reshape(length, P).to(torch.complex64).contiguous()
        flat = torch.view_as_real(t)                                   # [length, P, 2] fp32, interleaved
        out = torch.empty((P, ld), dtype=torch.float32, device=t.device)
        if ld > 2 * length:
            out[:, 2 * length:] = 0.0
        _lib.check(lib.csr_pack_planar(tptr(flat), _offset_ptr(flat, 4), 2 * P, 2, length, P, tptr(out), ld, stream_ptr()),
                   "csr_pack_planar")
    else:
        pix, P = pixel_shape(t, 2 * length)
        t = t.reshape(2 * length, P).to(torch.float32).contiguous()
        out = torch.empty((P, ld), dtype=torch.float32, device=t.device)
        if ld > 2 * length:
            out[:, 2 * length:] = 0.0
        _lib.check(lib.csr_pack_planar(tptr(t), _offset_ptr(t, 4 * length * P), P, 1, length, P, tptr(out), ld, stream_ptr()),
                   "csr_pack_planar")
    return out, pix


def pack_real(a, length, ld, device=None):
    """real ``(length, *pix)`` -> fp32 ``[P, ld]``."""
    t = to_device(a, device=device)
    pix, P = pixel_shape(t, length)
    t = t.reshape(length, P).to(torch.float32).contiguous()
    out = torch.empty((P, ld), dtype=torch.float32, device=t.device)
    if ld > length:
        out[:, length:] = 0.0
    _lib.check(_lib.load().csr_pack_planar(tptr(t), None, P, 1, length, P, tptr(out), ld, stream_ptr()), "csr_pack_planar")
    return out, pix


def unpack_complex(t, length, pix, like_torch=False):
    """fp32 ``[P, ld]`` planar -> complex64 ``(length, *pix)`` (numpy unless ``like_torch``)."""
    P = t.shape[0]
    c = torch.empty((length, P), dtype=torch.complex64, device=t.device)
    flat = torch.view_as_real(c)
    _lib.check(_lib.load().csr_unpack_planar(tptr(t), t.shape[1], length, P, tptr(flat), _offset_ptr(flat, 4), 2 * P, 2,
                                             stream_ptr()), "csr_unpack_planar")
    c = c.reshape((length,) + tuple(pix))
    return c if like_torch else c.cpu().numpy()


def unpack_complex_like(t, length, pix, like):
    """as ``unpack_complex`` but typed and placed like ``like``: numpy in -> numpy out, torch in -> torch on the
    same device (host tensors get the result copied back)."""
    if not is_torch(like):
        return unpack_complex(t, length, pix)
    return unpack_complex(t, length, pix, like_torch=True).to(like.device)


def unpack_real(t, length, pix, like_torch=False):
    P = t.shape[0]
    r = torch.empty((length, P), dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().csr_unpack_planar(tptr(t), t.shape[1], length, P, tptr(r), None, P, 1, stream_ptr()),
               "csr_unpack_planar")
    r = r.reshape((length,) + tuple(pix))
    return r if like_torch else r.cpu().numpy()


class SparseRows:
    """Compressed rows of a solution slab: for every line of sight the indices and values of its non-zero entries
    (``csr_sparse_count`` / ``csr_sparse_fill``).  Behaves like the ``(ncoef, *pix)`` array it stands for as far as
    ``len()``, ``shape`` and ``todense()`` go; ``offsets`` (P + 1, int64), ``indices`` (int32) and ``values`` (fp32) are
    torch tensors on the device the solution lives on, or NumPy arrays after ``cpu()``."""

    def __init__(self, offsets, indices, values, ncoef, pix):
        self.offsets, self.indices, self.values = offsets, indices, values
        self.ncoef, self.pix = int(ncoef), tuple(pix)

    @classmethod
    def from_planar(cls, x, ncoef, pix, capacity=None):
        """``x``: fp32 ``[P, ld]`` on the device.  ``capacity`` = None sizes the lists exactly (one 8-byte read-back, i.e. a
        synchronisation); an integer allocates that many entries without waiting for the device -- ``trim()`` later
        checks that the solution fitted."""
        lib = _lib.load()
        P = x.shape[0]
        counts = torch.empty((P,), dtype=torch.int32, device=x.device)
        _lib.check(lib.csr_sparse_count(tptr(x), x.shape[1], int(ncoef), P, tptr(counts), stream_ptr()), "csr_sparse_count")
        offsets = torch.zeros((P + 1,), dtype=torch.int64, device=x.device)
        torch.cumsum(counts, dim=0, out=offsets[1:])
        total = int(offsets[-1].item()) if capacity is None else int(capacity)
        indices = torch.empty((max(total, 1),), dtype=torch.int32, device=x.device)
        values = torch.empty((max(total, 1),), dtype=torch.float32, device=x.device)
        _lib.check(lib.csr_sparse_fill(tptr(x), x.shape[1], int(ncoef), P, tptr(offsets), tptr(indices), tptr(values), total,
                                       stream_ptr()), "csr_sparse_fill")
        rows = cls(offsets, indices[:total], values[:total], ncoef, pix)
        rows.lazy = capacity is not None
        return rows

    lazy = False

    def trim(self):
        """lists allocated with a ``capacity``: wait for the count, cut them to it (raises if they were too short)"""
        if self.lazy:
            total = int(self.offsets[-1])
            if total > self.values.shape[0]:
                raise CsrError("SparseRows: %d non-zeros do not fit the capacity of %d entries" % (total, self.values.shape[0]))
            self.indices, self.values, self.lazy = self.indices[:total], self.values[:total], False
        return self

    def __len__(self):
        return self.ncoef

    @property
    def shape(self):
        return (self.ncoef,) + self.pix

    @property
    def nnz(self):
        return int(self.trim().values.shape[0])

    @property
    def nbytes(self):
        return int(self.offsets.numel()) * 8 + self.nnz * 8

    def cpu(self):
        if not is_torch(self.offsets):
            return self
        self.trim()
        return SparseRows(self.offsets.cpu().numpy(), self.indices.cpu().numpy(), self.values.cpu().numpy(), self.ncoef, self.pix)

    def row(self, p):
        lo, hi = int(self.offsets[p]), int(self.offsets[p + 1])
        return self.indices[lo:hi], self.values[lo:hi]

    def todense(self):
        """NumPy ``(ncoef, *pix)`` fp32"""
        h = self.cpu()
        P = h.offsets.shape[0] - 1
        out = np.zeros((P, self.ncoef), dtype=np.float32)
        rows = np.repeat(np.arange(P), np.diff(h.offsets))
        out[rows, h.indices] = h.values
        return np.ascontiguousarray(out.T).reshape(self.shape)


_small_cache = {}


def small_to_device(arr, dtype, device):
    """Host vector (per-channel weights, spectral factors, per-pixel k / lambda / noise) -> read-only device tensor, cached by
    CONTENT.  A blocking host-to-device copy makes the host wait for everything queued on the stream before it: seven of them
    per reconstruction (the same few KB every slab) kept a streaming loop from queueing the next slab's work while the current
    one runs.  Arrays above 1 MB are not cached."""
    a = np.ascontiguousarray(arr)
    if a.nbytes > (1 << 20):
        return torch.as_tensor(a, dtype=dtype).to(device)
    key = (str(device), str(dtype), a.shape, a.dtype.str, hash(a.tobytes()))
    t = _small_cache.get(key)
    if t is None:
        if len(_small_cache) >= 256:
            _small_cache.pop(next(iter(_small_cache)))
        t = torch.as_tensor(a, dtype=dtype).to(device)
        _small_cache[key] = t
    return t


def per_pixel_vector(v, P, device, name):
    """scalar or (P,)/pixel-shaped array -> fp32 [P] on the device (read-only: small host arrays come from a content cache)."""
    if is_torch(v):
        t = v.to(device=device, dtype=torch.float32).reshape(-1)
    else:
        t = small_to_device(np.asarray(v, dtype=np.float64).reshape(-1), torch.float32, device)
    if t.numel() == 1:
        t = t.expand(P).contiguous()
    if t.numel() != P:
        raise ValueError("%s has %d entries, expected 1 or %d" % (name, t.numel(), P))
    return t.contiguous()


class Workspace:
    """grow-only scratch buffer"""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        return self.buf


class DevicePlan:
    """Owns a ``csr_plan`` (twiddle tables) for one (lambda2, l2_ref, phi) sampling on one GPU."""

    _cache = {}

    @classmethod
    def get(cls, lambda2, l2_ref, phi, device=None):
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        l2 = np.ascontiguousarray(lambda2, dtype=np.float64)
        ph = np.ascontiguousarray(phi, dtype=np.float64)
        key = (hashlib.sha1(l2.tobytes()).hexdigest(), hashlib.sha1(ph.tobytes()).hexdigest(), float(l2_ref),
               dev.index)
        plan = cls._cache.get(key)
        if plan is None:
            if len(cls._cache) >= 4:  # tables are large (64 MB .. 4 GB): keep only a few
                cls._cache.pop(next(iter(cls._cache))).close()
            plan = cls(l2, float(l2_ref), ph, dev)
            cls._cache[key] = plan
        return plan

    def __init__(self, lambda2, l2_ref, phi, device):
        self.lib = _lib.load()
        self.device = device
        self.m, self.n = int(lambda2.shape[0]), int(phi.shape[0])
        handle = ctypes.c_void_p()
        _lib.check(self.lib.csr_plan_create(ctypes.byref(handle), _lib.dptr(lambda2), self.m, l2_ref, _lib.dptr(phi),
                                            self.n, device.index), "csr_plan_create")
        self.handle = handle
        m_, n_, ld2m, ld2n, tb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _lib.check(self.lib.csr_plan_dims(handle, m_, n_, ld2m, ld2n, tb), "csr_plan_dims")
        self.ld2m, self.ld2n, self.table_bytes = ld2m.value, ld2n.value, tb.value
        self.ws = Workspace()

    def close(self):
        if self.handle:
            self.lib.csr_plan_destroy(self.handle)
            self.handle = None

    def _transform(self, fn, name, inp, ld_out, chan_scale, row_scale):
        P = inp.shape[0]
        out = torch.empty((P, ld_out), dtype=torch.float32, device=self.device)
        need = self.lib.csr_transform_ws_bytes(self.handle, P)
        ws = self.ws.get(need, self.device)
        per_pixel = 0
        if chan_scale is not None:
            chan_scale = chan_scale.contiguous()
            per_pixel = int(chan_scale.dim() == 2)
        _lib.check(fn(self.handle, tptr(inp), tptr(chan_scale), per_pixel, tptr(row_scale), tptr(out), P, tptr(ws),
                      ws.numel(), stream_ptr()), name)
        return out

    def forward(self, x, chan_scale=None, row_scale=None):
        """[P, ld2n] -> [P, ld2m]; out = row_scale * chan_scale * (E x)"""
        return self._transform(self.lib.csr_forward, "csr_forward", x, self.ld2m, chan_scale, row_scale)

    def adjoint(self, b, chan_scale=None, row_scale=None):
        """[P, ld2m] -> [P, ld2n]; out = row_scale * E^H (chan_scale * b)"""
        return self._transform(self.lib.csr_adjoint, "csr_adjoint", b, self.ld2n, chan_scale, row_scale)

    def derotate(self, data, phi_gal):
        _lib.check(self.lib.csr_derotate(self.handle, tptr(data), tptr(phi_gal), data.shape[0], stream_ptr()),
                   "csr_derotate")

    def fista(self, data, w, s, k, x, lambda0, noise, iters, step=1.0, norm_factor=1.0, wavelet=None,
              use_dirty_as_guess=False, want_dirty=True, want_model=True, want_delta=False, iter_cap=None,
              want_mma_blocks=False, out=None, f_dirty_in=None):
        """Run the fused loop.  All tensors are fp32 on the device; ``x`` ([P, ldx]) is updated in place.
        Returns dict(f_dirty, model, residual, cost[P,2], delta[, mma_blocks]).  ``mma_blocks`` (measurement aid) is the
        number of 128 x 256 x 64 tensor-core blocks the in-loop GEMMs issued, by kind: [three-product, fp16 screening,
        fp8 screening, unused] (3 per tile and K block when dense).  ``out``: the dict a previous call with the same shapes
        and ``want_*`` flags returned -- its tensors are written again instead of new ones being allocated (a slab loop that
        reuses its buffers repeats the call with identical arguments, which the library replays as one CUDA graph).
        ``f_dirty_in``: planar ``[P, ld2n]`` dirty spectrum of these data and weights when the caller already has it
        (``plan.adjoint``): the loop does not compute it again; with ``want_dirty=False`` it is returned as ``f_dirty``."""
        P = data.shape[0]
        dev = self.device
        if out is not None:
            for key, want in (("f_dirty", want_dirty), ("model", want_model), ("residual", want_model), ("cost", want_model),
                              ("delta", want_delta), ("mma_blocks", want_mma_blocks)):
                t = out.get(key)
                if (t is not None) != bool(want) or (t is not None and t.shape[0] != (16 if key == "mma_blocks" else P)):
                    raise ValueError("out does not match this call (%s)" % key)
            if want_model:
                out["residual"].zero_()
            if want_delta:
                out["delta"].zero_()
            if want_mma_blocks:
                out["mma_blocks"].zero_()
        else:
            out = dict(f_dirty=None, model=None, residual=None, cost=None, delta=None)
            if want_dirty:
                out["f_dirty"] = torch.empty((P, self.ld2n), dtype=torch.float32, device=dev)
            if want_model:
                out["model"] = torch.empty((P, self.ld2m), dtype=torch.float32, device=dev)
                out["residual"] = torch.zeros((P, self.ld2m), dtype=torch.float32, device=dev)
                out["cost"] = torch.empty((P, 2), dtype=torch.float32, device=dev)
            if want_delta:
                out["delta"] = torch.zeros((P,), dtype=torch.float32, device=dev)
            if want_mma_blocks:
                out["mma_blocks"] = torch.zeros((16,), dtype=torch.int64, device=dev)
        whandle = wavelet.handle if wavelet is not None else None
        need = self.lib.csr_fista_ws_bytes(self.handle, whandle, P)
        ws = self.ws.get(need, dev)
        a = _lib.FistaArgs()
        a.mma_blocks = out["mma_blocks"].data_ptr() if want_mma_blocks else None
        a.P, a.iters, a.step, a.norm_factor = P, int(iters), float(step), float(norm_factor)
        if s.dim() == 2 and w.dim() == 1:       # per-pixel spectral factors need the per-pixel channel diagonals
            w = w[None, :].expand(P, self.m).contiguous()
        a.w_per_pixel = int(w.dim() == 2)
        a.s_per_pixel = int(s.dim() == 2)       # per-pixel spectral factors [P, m] (prepare_kernel reads the row of p)
        for name, t, pp in (("w", w, a.w_per_pixel), ("s", s, a.s_per_pixel)):
            want = (P, self.m) if pp else (self.m,)
            if tuple(t.shape) != want:
                raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), want))
        a.data, a.w, a.s, a.k = data.data_ptr(), w.data_ptr(), s.data_ptr(), k.data_ptr()
        a.x, a.ldx = x.data_ptr(), x.shape[1]
        a.lambda0, a.noise = lambda0.data_ptr(), noise.data_ptr()
        a.iter_cap = iter_cap.data_ptr() if iter_cap is not None else None
        for key in ("f_dirty", "model", "residual", "cost", "delta"):
            setattr(a, key, out[key].data_ptr() if out[key] is not None else None)
        a.use_dirty_as_guess = int(bool(use_dirty_as_guess))
        if f_dirty_in is not None:
            if tuple(f_dirty_in.shape) != (P, self.ld2n) or f_dirty_in.dtype != torch.float32 or not f_dirty_in.is_contiguous():
                raise ValueError("f_dirty_in must be a contiguous fp32 [P, ld2n] tensor")
            a.f_dirty_in = f_dirty_in.data_ptr()
            if out["f_dirty"] is None:
                out["f_dirty"] = f_dirty_in
        _lib.check(self.lib.csr_fista(self.handle, whandle, ctypes.byref(a), tptr(ws), ws.numel(), stream_ptr()),
                   "csr_fista")
        return out


class DeviceStack:
    """Per-line-of-sight samplings (kernel K4, ``csr_nudft_*`` / ``csr_fista_nu``): lambda2 - l2_ref of every line of
    sight as an fp64 ``[P, m]`` device matrix (zero padded past ``m_count[p]``) on a uniform phi grid.  No tables."""

    def __init__(self, a, m_count, phi, device=None):
        require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        phi = np.asarray(phi, dtype=np.float64)
        if phi.ndim != 1 or phi.size < 2:
            raise ValueError("per-pixel operator needs a 1-D phi grid")
        dphi = (phi[-1] - phi[0]) / (phi.size - 1)
        if not np.allclose(np.diff(phi), dphi, rtol=1e-9, atol=0.0):
            raise ValueError("per-pixel operator needs a uniform phi grid (Parameter.calculate_cellsize builds one)")
        self.phi0, self.dphi = float(phi[0]), float(dphi)
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.P, self.m = int(a.shape[0]), int(a.shape[1])
        self.n = int(phi.size)
        self.ld2m, self.ld2n = (2 * self.m + 7) // 8 * 8, (2 * self.n + 7) // 8 * 8
        self.a = torch.as_tensor(a).to(self.device)
        self.m_count = torch.as_tensor(np.ascontiguousarray(m_count, dtype=np.int32)).to(self.device)
        self.ws = Workspace()

    def _sampling(self):
        s = _lib.NuSampling()
        s.lambda2, s.ldl, s.m_count = self.a.data_ptr(), self.m, self.m_count.data_ptr()
        s.m, s.n, s.phi0, s.dphi = self.m, self.n, self.phi0, self.dphi
        return s

    def forward(self, x, chan_scale=None, row_scale=None):
        """[P, ld2n] -> [P, ld2m]"""
        out = torch.zeros((self.P, self.ld2m), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.csr_nudft_forward(ctypes.byref(self._sampling()), tptr(x), x.shape[1], tptr(chan_scale),
                                              tptr(row_scale), tptr(out), self.ld2m, self.P, stream_ptr()),
                   "csr_nudft_forward")
        return out

    def adjoint(self, b, chan_scale=None, row_scale=None):
        """[P, ld2m] -> [P, ld2n]"""
        out = torch.zeros((self.P, self.ld2n), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.csr_nudft_adjoint(ctypes.byref(self._sampling()), tptr(b), b.shape[1], tptr(chan_scale),
                                              tptr(row_scale), tptr(out), self.ld2n, self.P, stream_ptr()),
                   "csr_nudft_adjoint")
        return out

    def fista(self, data, w, s, k, x, lambda0, noise, iters, step=1.0, norm_factor=1.0, wavelet=None,
              use_dirty_as_guess=False, want_dirty=True, want_model=True, want_delta=False, iter_cap=None):
        """same contract as ``DevicePlan.fista``; ``w`` and ``s`` are ``[P, m]``"""
        P, dev = self.P, self.device
        out = dict(f_dirty=None, model=None, residual=None, cost=None, delta=None)
        if want_dirty:
            out["f_dirty"] = torch.empty((P, self.ld2n), dtype=torch.float32, device=dev)
        if want_model:
            out["model"] = torch.empty((P, self.ld2m), dtype=torch.float32, device=dev)
            out["residual"] = torch.zeros((P, self.ld2m), dtype=torch.float32, device=dev)
            out["cost"] = torch.empty((P, 2), dtype=torch.float32, device=dev)
        if want_delta:
            out["delta"] = torch.zeros((P,), dtype=torch.float32, device=dev)
        whandle = wavelet.handle if wavelet is not None else None
        samp = self._sampling()
        need = self.lib.csr_fista_nu_ws_bytes(ctypes.byref(samp), whandle, P)
        ws = self.ws.get(need, dev)
        a = _lib.FistaArgs()
        a.P, a.iters, a.step, a.norm_factor = P, int(iters), float(step), float(norm_factor)
        a.w_per_pixel = a.s_per_pixel = 1
        a.data, a.w, a.s, a.k = data.data_ptr(), w.data_ptr(), s.data_ptr(), k.data_ptr()
        a.x, a.ldx = x.data_ptr(), x.shape[1]
        a.lambda0, a.noise = lambda0.data_ptr(), noise.data_ptr()
        a.iter_cap = iter_cap.data_ptr() if iter_cap is not None else None
        for key in ("f_dirty", "model", "residual", "cost", "delta"):
            setattr(a, key, out[key].data_ptr() if out[key] is not None else None)
        a.use_dirty_as_guess = int(bool(use_dirty_as_guess))
        _lib.check(self.lib.csr_fista_nu(ctypes.byref(samp), whandle, ctypes.byref(a), tptr(ws), ws.numel(),
                                         stream_ptr()), "csr_fista_nu")
        return out


class DeviceWavelet:
    """Owns a ``csr_wavelet`` for one (filter bank, level, N) configuration."""

    def __init__(self, kind, bank, level, N, append_signal, norm, device=None):
        require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        fs = [np.ascontiguousarray(f, dtype=np.float64) for f in bank]
        F = fs[0].shape[0]
        handle = ctypes.c_void_p()
        _lib.check(self.lib.csr_wavelet_create(ctypes.byref(handle), int(kind), _lib.dptr(fs[0]), _lib.dptr(fs[1]),
                                               _lib.dptr(fs[2]), _lib.dptr(fs[3]), F, int(level), int(N),
                                               int(bool(append_signal)), int(bool(norm)), self.device.index),
                   "csr_wavelet_create")
        self.handle = handle
        self.N = int(N)
        self.ncoef = self.lib.csr_wavelet_ncoef(handle)
        self.levels = self.lib.csr_wavelet_levels(handle)
        sizes = (ctypes.c_int * 40)()
        nb = self.lib.csr_wavelet_level_sizes(handle, sizes, 40)
        self.sizes = [sizes[i] for i in range(nb)]
        self.ldc = (self.ncoef + 7) // 8 * 8
        self.ldn = (self.N + 7) // 8 * 8
        self.ws = Workspace()

    def __del__(self):
        try:
            if self.handle:
                self.lib.csr_wavelet_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def decompose(self, signal):
        """[P, ld >= N] -> [P, ldc]"""
        P = signal.shape[0]
        coef = torch.zeros((P, self.ldc), dtype=torch.float32, device=self.device)
        ws = self.ws.get(self.lib.csr_wavelet_ws_bytes(self.handle, P), self.device)
        _lib.check(self.lib.csr_wavelet_decompose(self.handle, tptr(signal), signal.shape[1], tptr(coef), self.ldc, P,
                                                  tptr(ws), ws.numel(), stream_ptr()), "csr_wavelet_decompose")
        return coef

    def reconstruct(self, coef):
        """[P, ld >= ncoef] -> [P, ldn]"""
        P = coef.shape[0]
        sig = torch.zeros((P, self.ldn), dtype=torch.float32, device=self.device)
        ws = self.ws.get(self.lib.csr_wavelet_ws_bytes(self.handle, P), self.device)
        _lib.check(self.lib.csr_wavelet_reconstruct(self.handle, tptr(coef), coef.shape[1], tptr(sig), self.ldn, P,
                                                    tptr(ws), ws.numel(), stream_ptr()), "csr_wavelet_reconstruct")
        return sig


# ---- restore-step helpers (csrc/restore.cu) -----------------------------------------------------------------
def gaussian_beam_taps(rmtf_fwhm, cellsize):
    """The restore beam of ``Parameter.convolve`` (reconstruction/parameter.py:139-156): sigma in whole phi cells,
    ``astropy.convolution.Gaussian1DKernel(stddev)`` = Gaussian sampled at integer offsets over an odd window of
    about 8 sigma, normalised to unit sum.  (astropy is not a dependency here; parity of this kernel is unpinned.)"""
    sigma_pix = int(np.round(rmtf_fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0))) / cellsize))
    if sigma_pix < 1:
        return np.ones(1, dtype=np.float64), sigma_pix
    size = int(np.ceil(8 * sigma_pix))
    size += 1 - size % 2
    x = np.arange(size) - size // 2
    taps = np.exp(-0.5 * (x / float(sigma_pix)) ** 2) / (np.sqrt(2 * np.pi) * sigma_pix)
    return taps / taps.sum(), sigma_pix


def convolve_phi(x, n, taps, add=None):
    """[P, ld >= 2n] planar -> same shape: beam * x (+ add)"""
    lib = _lib.load()
    out = torch.zeros_like(x)
    t = np.ascontiguousarray(taps, dtype=np.float32)
    _lib.check(lib.csr_convolve_phi(tptr(x), x.shape[1], int(n), x.shape[0], t.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                    len(t), tptr(add), tptr(out), stream_ptr()), "csr_convolve_phi")
    return out


def edge_noise(F, n, edge_mask, eta=1.0):
    """[P, ld] planar dirty spectra -> [P] noise estimate from the masked (edge) phi cells"""
    lib = _lib.load()
    noise = torch.empty((F.shape[0],), dtype=torch.float32, device=F.device)
    _lib.check(lib.csr_edge_noise(tptr(F), F.shape[1], int(n), F.shape[0], tptr(edge_mask), float(eta), tptr(noise),
                                  stream_ptr()), "csr_edge_noise")
    return noise


def peak_stats(F, n):
    """[P, ld] planar -> (argmax [P] int32, stats [P, 3] = |F| at peak, parabola offset, parabola peak)"""
    lib = _lib.load()
    idx = torch.empty((F.shape[0],), dtype=torch.int32, device=F.device)
    stats = torch.empty((F.shape[0], 3), dtype=torch.float32, device=F.device)
    _lib.check(lib.csr_peak(tptr(F), F.shape[1], int(n), F.shape[0], tptr(idx), tptr(stats), stream_ptr()), "csr_peak")
    return idx, stats
