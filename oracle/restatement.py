"""Oracle B -- float64 NumPy restatement of the CS-ROMER reconstruction hot path.

TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, never the thing shipped or measured.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; nothing under ``csromer_b200/`` does.

Parity status
  * exact-DFT operator, Chi2 gradient, L1 prox, FISTA loop: PINNED against the reference's own classes
    (Oracle A, ``oracle/ref_shim.py``) through the golden fixtures in ``tests/golden/`` made by
    ``oracle/make_golden.py`` (agreement 3e-8, see tests/test_oracle_golden.py).
  * wavelet transforms (pywt 1.4.1 semantics): PARITY UNPINNED -- PyWavelets is a third-party dependency
    (``requirements.txt:12``) absent from ``/root/reference`` and from this image; the algorithm published by
    PyWavelets is restated here and pinned only by pywt-documentation known answers, perfect reconstruction
    and Parseval identities (tests/test_oracle_wavelets.py).
  * ``NUFFT1D`` arithmetic (pynufft 2023.1.1, ``requirements.txt:11``): PARITY UNPINNED and deliberately not
    reproduced; the exact operator is used instead.

All arrays are frequency/phi FIRST and pixel LAST: ``data (m,)`` or ``(m, P)``, ``x (n,)`` or ``(n, P)``,
which is the batch extension of the reference's 1-D layout (``README.md:327`` slices ``cube[:, i, j]``).
Citations are into /root/reference/src/csromer/.
"""
import numpy as np

C_LIGHT = 299792458.0  # scipy.constants.c, used by base/dataset.py:14


# ----------------------------------------------------------------------------------------------------------
# containers: Dataset scalars (base/dataset.py) and Parameter grid (reconstruction/parameter.py)
# ----------------------------------------------------------------------------------------------------------
def lambda2_from_nu(nu):
    """base/dataset.py:318-320 ``nu_to_l2``: lambda2 = (c/nu)^2 reversed to ascending order."""
    nu = np.asarray(nu, dtype=np.float64)
    return ((C_LIGHT / nu) ** 2)[::-1]


def spectral_factor(lambda2, spectral_idx):
    """base/dataset.py:207-218: nu re-derived from lambda2, nu_0 mid-band, s = (nu/nu_0)^alpha."""
    nu = C_LIGHT / np.sqrt(lambda2)
    nu_0 = 0.5 * (nu.min() + nu.max())
    return (nu / nu_0) ** spectral_idx, nu_0


def dataset_scalars(lambda2, sigma=None, spectral_idx=0.0, l2_ref=0.0):
    """w = 1/sigma^2 (dataset.py:278-282), s, k = sum(w/s) (dataset.py:178-182), theo_noise
    (dataset.py:362-372), mean |delta lambda2| over w>0 channels (dataset.py:351-360)."""
    lambda2 = np.asarray(lambda2, dtype=np.float64)
    m = lambda2.shape[0]
    w = np.ones(m) if sigma is None else 1.0 / np.asarray(sigma, dtype=np.float64) ** 2
    s, nu_0 = spectral_factor(lambda2, spectral_idx)
    s_b = s if w.ndim == 1 else s[:, None]
    k = np.sum(w / s_b, axis=0)
    theo_noise = None if np.all(w == 1.0) else 1.0 / np.sqrt(np.sum(w, axis=0))
    return dict(w=w, s=s, k=k, nu_0=nu_0, theo_noise=theo_noise, l2_ref=l2_ref, m=m)


def phi_grid(lambda2, w=None, oversampling=8.0, cellsize=None, set_size_pow_2=False):
    """reconstruction/parameter.py:62-116 ``calculate_cellsize``."""
    lambda2 = np.asarray(lambda2, dtype=np.float64)
    l2_min = np.min(lambda2[np.nonzero(lambda2)])
    l2_max = np.max(lambda2)
    sel = lambda2 if w is None else lambda2[np.asarray(w) > 0.0]
    delta_l2_mean = np.mean(np.abs(np.diff(sel)))
    fwhm = 2.0 * np.sqrt(3.0) / (l2_max - l2_min)
    theo = np.pi / l2_min
    phi_max = max(np.sqrt(3) / delta_l2_mean, fwhm * 10.0)
    phi_r = min(fwhm, theo) / oversampling if cellsize is None else cellsize
    temp = int(np.int32(np.floor(2 * phi_max / phi_r)))
    if set_size_pow_2:
        n = 1 if temp == 0 else (temp if temp & (temp - 1) == 0 else 1 << temp.bit_length())
    else:
        n = int(temp - temp % 32)
    cell = 2 * phi_max / n
    phi = cell * np.arange(-(n / 2), (n / 2), 1)
    return dict(phi=phi, n=n, cellsize=cell, rmtf_fwhm=fwhm, max_recovered_width=theo, max_faraday_depth=phi_max)


def complex_to_real(z):
    """utils/utilities.py:31-32 -- planar [Re | Im] along axis 0."""
    return np.concatenate([z.real, z.imag], axis=0)


def real_to_complex(v):
    """utils/utilities.py:27-28."""
    h = v.shape[0] // 2
    return v[:h] + 1j * v[h:]


# ----------------------------------------------------------------------------------------------------------
# exact operator (transformers/dfts/ndft.py) with the twiddle matrix cached
# ----------------------------------------------------------------------------------------------------------
class ExactDFT:
    """E[i, j] = exp(+2i (lambda2_i - l2_ref) phi_j), built once in complex128 (ndft.py:19 rebuilds it on
    every call; values are identical).  ``w`` may be (m,) or (m, P) (per-pixel weights / flags); ``k`` follows.
    ``k`` may also be passed explicitly because in-place flagging leaves it stale in the reference
    (transformers/flaggers/mean_flagger.py:42)."""

    def __init__(self, lambda2, phi, w=None, s=None, k=None, l2_ref=0.0, rebuild_twiddles=False, cache_adjoint=False):
        # rebuild_twiddles=True reproduces the reference's COST profile as well as its values: ndft.py:19,25,34,42
        # call np.exp on the full m x n matrix inside every transform.  Used only by bench.py's CPU baseline.
        self.rebuild_twiddles = rebuild_twiddles
        # cache_adjoint=True also keeps conj(E).T (bench.py's in-run parity sample at the large samplings: the
        # conjugate copy of a 2 GB matrix per call would dominate the check); values are identical
        self.cache_adjoint = cache_adjoint and not rebuild_twiddles
        self._Eh = None
        self.lambda2 = np.asarray(lambda2, dtype=np.float64)
        self.phi = np.asarray(phi, dtype=np.float64)
        self.m, self.n = self.lambda2.shape[0], self.phi.shape[0]
        self.w = np.ones(self.m) if w is None else np.asarray(w, dtype=np.float64)
        self.s = np.ones(self.m) if s is None else np.asarray(s, dtype=np.float64)
        s_b = self.s if self.w.ndim == 1 else self.s[:, None]
        self.k = np.sum(self.w / s_b, axis=0) if k is None else k
        self._l2 = self.lambda2 - l2_ref
        self._E = None if rebuild_twiddles else np.exp(2.0j * self._l2[:, None] * self.phi[None, :])  # (m, n)

    @property
    def E(self):
        if self._E is not None:
            return self._E
        return np.exp(2.0j * self._l2[:, None] * self.phi[None, :])

    def _col(self, v, like):
        return v[:, None] if (v.ndim == 1 and like.ndim == 2) else v

    def forward(self, x):
        """ndft.py:17-19: P_i = sum_j x_j exp(+2i..)."""
        return self.E @ x

    def forward_normalized(self, x):
        """ndft.py:21-28 as intended (``self.k`` -> ``dataset.k``, nufft.py:54-59); w == 0 channels give 0
        (the reference leaves them uninitialised)."""
        b = self.E @ (x * self.k)
        w = self._col(self.w, b)
        b = np.where(w > 0.0, b / np.where(w > 0.0, w, 1.0), 0.0)
        return b * self._col(self.s, b) / self.n

    def backward(self, b):
        """ndft.py:30-36: F_j = (1/k) sum_i (w_i/s_i) b_i exp(-2i..)."""
        wt = self._col(self.w, b) / self._col(self.s, b)
        if self.rebuild_twiddles:  # the reference builds the conjugate matrix directly (ndft.py:34)
            return (np.exp(-2.0j * self.phi[:, None] * self._l2[None, :]) @ (wt * b)) / self.k
        if self.cache_adjoint:
            if self._Eh is None:
                self._Eh = np.ascontiguousarray(self.E.conj().T)
            return (self._Eh @ (wt * b)) / self.k
        return (self.E.conj().T @ (wt * b)) / self.k

    def rmtf(self):
        """ndft.py:38-43."""
        wt = self.w / self._col(self.s, self.w)
        return (self.E.conj().T @ wt) / self.k


# ----------------------------------------------------------------------------------------------------------
# wavelets: pywt 1.4.1 semantics restated
# ----------------------------------------------------------------------------------------------------------
_REC_LO_FROM_DEC = {
    "db1": [0.7071067811865476, 0.7071067811865476],
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313,
            0.8068915093133388, 0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523],
    "coif1": [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
              0.3378976624578092, -0.0727326195128539],
    "coif2": [-0.0007205494453645122, -0.0018232088707029932, 0.0056114348193944995, 0.023680171946334084,
              -0.0594344186464569, -0.0764885990783064, 0.41700518442169254, 0.8127236354455423,
              0.3861100668211622, -0.06737255472196302, -0.04146493678175915, 0.016387336463522112],
    "sym4": [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
             0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
}
_REC_LO_FROM_DEC["haar"] = _REC_LO_FROM_DEC["db1"]
_REC_LO_FROM_DEC["sym2"] = _REC_LO_FROM_DEC["db2"]
_REC_LO_FROM_DEC["sym3"] = _REC_LO_FROM_DEC["db3"]


def filter_bank(name):
    """(dec_lo, dec_hi, rec_lo, rec_hi) in pywt orientation.  Orthogonal families follow the quadrature-mirror
    rule rec_lo[t] = dec_lo[F-1-t], dec_hi[t] = (-1)^(t+1) rec_lo[t], rec_hi[t] = (-1)^t dec_lo[t].
    'IUWT' is the reference's own bank (dictionaries/undecimated.py:20-25: g in the low-pass slot)."""
    if name == "IUWT":
        h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
        g = [-1 / 16, -4 / 16, 10 / 16, -4 / 16, -1 / 16]
        d = [0.0, 0.0, 1.0, 0.0, 0.0]
        return tuple(np.array(f + [0.0]) for f in (g, h, d, d))  # pywt pads odd banks to even length
    if name == "bior2.2":
        a, b, c, r = 0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.7071067811865476
        return (np.array([0, -a, b, c, b, -a]), np.array([0, b, -r, b, 0, 0]),
                np.array([0, b, r, b, 0, 0]), np.array([0, a, b, -c, b, a]))
    if name not in _REC_LO_FROM_DEC:
        raise NotImplementedError("The wavelet has not been implemented: %s" % name)
    dec_lo = np.array(_REC_LO_FROM_DEC[name], dtype=np.float64)
    F = dec_lo.shape[0]
    rec_lo = dec_lo[::-1].copy()
    sign = np.array([(-1.0) ** (t + 1) for t in range(F)])
    dec_hi = sign * rec_lo
    rec_hi = -sign * dec_lo
    return dec_lo, dec_hi, rec_lo, rec_hi


def swt_max_level(N):
    """pywt.swt_max_level: number of times N divides by two."""
    lev = 0
    while N > 0 and N % 2 == 0:
        N //= 2
        lev += 1
    return lev


def dwt_max_level(N, F):
    """pywt.dwt_max_level = floor(log2(N / (F - 1)))."""
    if F < 2 or N < F - 1:
        return 0
    return int(np.floor(np.log2(N / (F - 1.0))))


def _swt_level(a, f, step):
    """one undecimated level along axis 0: c[o] = sum_t f[t] a[(o + (F*step)//2 - t*step) mod N]."""
    F = f.shape[0]
    out = np.zeros_like(a)
    shift0 = (F * step) // 2
    for t in range(F):
        if f[t] != 0.0:
            out += f[t] * np.roll(a, -(shift0 - t * step), axis=0)
    return out


def swt(x, bank, level=None, norm=False):
    """pywt.swt(trim_approx=True) along axis 0 -> [cA_L, cD_L, ..., cD_1] (dictionaries/undecimated.py:75-81)."""
    dec_lo, dec_hi = bank[0], bank[1]
    if norm:
        dec_lo, dec_hi = dec_lo / np.sqrt(2), dec_hi / np.sqrt(2)
    N = x.shape[0]
    L = swt_max_level(N) if level is None else level
    a = np.asarray(x, dtype=np.float64)
    details = []
    for lev in range(1, L + 1):
        step = 1 << (lev - 1)
        details.append(_swt_level(a, dec_hi, step))
        a = _swt_level(a, dec_lo, step)
    return [a] + details[::-1]


def _idwt_per(cA, cD, rec_lo, rec_hi):
    """single-level periodization IDWT along axis 0: x[(2o + t - F/2 + 1) mod N] += cA[o] rec_lo[t] + cD[o] rec_hi[t]."""
    F = rec_lo.shape[0]
    half = cA.shape[0]
    N = 2 * half
    x = np.zeros((N,) + cA.shape[1:], dtype=np.float64)
    o = np.arange(half)
    for t in range(F):
        idx = (2 * o + t - F // 2 + 1) % N
        contrib = cA * rec_lo[t] + cD * rec_hi[t]
        np.add.at(x, idx, contrib)
    return x


def iswt(coeffs, bank, norm=False):
    """pywt.iswt for the trim_approx list [cA_L, cD_L, ..., cD_1] (dictionaries/undecimated.py:159)."""
    rec_lo, rec_hi = bank[2], bank[3]
    if norm:  # pywt.iswt(norm=True) rescales the bank by sqrt(2) (swt(norm=True) by 1/sqrt(2))
        rec_lo, rec_hi = rec_lo * np.sqrt(2), rec_hi * np.sqrt(2)
    out = np.array(coeffs[0], dtype=np.float64, copy=True)
    L = len(coeffs) - 1
    N = out.shape[0]
    for j in range(L, 0, -1):
        step = 1 << (j - 1)
        cD = coeffs[L - j + 1]
        for first in range(step):
            idx = np.arange(first, N, step)
            ev, od = idx[0::2], idx[1::2]
            x1 = _idwt_per(out[ev], cD[ev], rec_lo, rec_hi)
            x2 = _idwt_per(out[od], cD[od], rec_lo, rec_hi)
            x2 = np.roll(x2, 1, axis=0)
            out[idx] = (x1 + x2) / 2.0
    return out


def iswt_fast(coeffs, bank, norm=False):
    """``iswt`` in closed form: pywt's average of the even- and odd-phase periodization IDWTs of one a-trous level
    collapses to ONE gather, a'[o] = 1/2 sum_t (rec_lo[t] a[.] + rec_hi[t] d[.]) at (o + step (F/2 - 1 - t)) mod N.
    Same values as ``iswt`` (the pywt-literal double loop; tests/test_oracle_wavelets.py holds them together) at a
    fraction of the cost -- used where the literal form would take minutes (bench.py's in-run parity sample)."""
    rec_lo, rec_hi = bank[2], bank[3]
    if norm:
        rec_lo, rec_hi = rec_lo * np.sqrt(2), rec_hi * np.sqrt(2)
    out = np.array(coeffs[0], dtype=np.float64, copy=True)
    L = len(coeffs) - 1
    F = rec_lo.shape[0]
    for j in range(L, 0, -1):
        step = 1 << (j - 1)
        cD = coeffs[L - j + 1]
        acc = np.zeros_like(out)
        for t in range(F):
            sh = step * (F // 2 - 1 - t)
            if rec_lo[t] != 0.0:
                acc += rec_lo[t] * np.roll(out, -sh, axis=0)
            if rec_hi[t] != 0.0:
                acc += rec_hi[t] * np.roll(cD, -sh, axis=0)
        out = 0.5 * acc
    return out


def _dwt_per(x, f_lo, f_hi):
    """single-level periodization DWT along axis 0; odd lengths repeat the last sample once."""
    N = x.shape[0]
    if N % 2:
        x = np.concatenate([x, x[-1:]], axis=0)
        N += 1
    F = f_lo.shape[0]
    o = np.arange(N // 2)
    cA = np.zeros((N // 2,) + x.shape[1:], dtype=np.float64)
    cD = np.zeros_like(cA)
    for t in range(F):
        src = x[(2 * o + F // 2 - t) % N]
        cA += f_lo[t] * src
        cD += f_hi[t] * src
    return cA, cD


def wavedec_per(x, bank, level=None):
    """pywt.wavedec(mode='periodization') along axis 0 -> [cA_L, cD_L, ..., cD_1] (dictionaries/discrete.py:37-39)."""
    dec_lo, dec_hi = bank[0], bank[1]
    L = dwt_max_level(x.shape[0], dec_lo.shape[0]) if level is None else level
    a = np.asarray(x, dtype=np.float64)
    details = []
    for _ in range(L):
        a, d = _dwt_per(a, dec_lo, dec_hi)
        details.append(d)
    return [a] + details[::-1]


def waverec_per(coeffs, bank):
    """pywt.waverec(mode='periodization') (dictionaries/discrete.py:80-89)."""
    rec_lo, rec_hi = bank[2], bank[3]
    a = coeffs[0]
    for d in coeffs[1:]:
        if a.shape[0] == d.shape[0] + 1:
            a = a[:-1]
        a = _idwt_per(a, d, rec_lo, rec_hi)
    return a


# ---- decimated transform with the other PyWavelets signal-extension modes ------------------------------------
# dictionaries/discrete.py:37-39,80-89 hand ``self.mode`` to pywt.wavedec / pywt.waverec.  PyWavelets (a dependency that
# is not vendored in the reference; pyproject pins no version, algorithm as of PyWavelets 1.x, pywt/_extensions/c/
# convolution.template.c ``downsampling_convolution`` / ``upsampling_convolution_valid_sf``) computes, for every
# non-periodization mode, the FULL convolution of the extended signal with the decomposition filter sampled at the odd
# positions,  cA[k] = sum_j dec_lo[j] x_ext[2k + 1 - j],  k = 0 .. (N + F - 1)//2 - 1,  and reconstructs with the part of
# the up-sampled convolution that does not depend on the extension ("valid"), 2 Nc - F + 2 samples.
# Pinned: the PyWavelets documentation value of pywt.dwt([3,7,1,1,-2,5,4,6], 'db2') (default mode 'symmetric') in
# tests/test_oracle_wavelets.py; the extensions themselves against numpy.pad (the correspondence pywt.pad documents).
EXT_MODES = ("zero", "constant", "symmetric", "reflect", "periodic", "smooth", "antisymmetric", "antireflect")


def ext_sample(x, i, mode):
    """value of the extended signal at (array of) integer positions ``i`` (any distance outside [0, N)); axis 0 of x."""
    x = np.asarray(x, dtype=np.float64)
    N = x.shape[0]
    i = np.asarray(i, dtype=np.int64)
    inside = (i >= 0) & (i < N)
    shape = (-1,) + (1,) * (x.ndim - 1)
    if mode == "zero":
        out = x[np.clip(i, 0, N - 1)] * inside.reshape(shape)
    elif mode == "constant":
        out = x[np.clip(i, 0, N - 1)]
    elif mode == "periodic":
        out = x[i % N]
    elif mode in ("symmetric", "antisymmetric"):
        j = i % (2 * N)
        back = j >= N
        out = x[np.where(back, 2 * N - 1 - j, j)]
        if mode == "antisymmetric":
            out = out * np.where(back, -1.0, 1.0).reshape(shape)
    elif mode == "reflect":
        if N == 1:
            out = x[np.zeros_like(i)]
        else:
            p = 2 * N - 2
            j = i % p
            out = x[np.where(j >= N, p - j, j)]
    elif mode == "antireflect":
        if N == 1:
            out = x[np.zeros_like(i)]
        else:
            p = 2 * N - 2
            cyc = np.floor_divide(i, p)
            j = i - cyc * p
            back = j >= N
            base = np.where(back.reshape(shape), 2.0 * x[N - 1] - x[np.where(back, p - j, 0)], x[np.where(back, 0, j)])
            out = base + 2.0 * cyc.reshape(shape) * (x[N - 1] - x[0])
    elif mode == "smooth":
        if N == 1:
            out = x[np.zeros_like(i)]
        else:
            left = x[0] + i.reshape(shape) * (x[1] - x[0])
            right = x[N - 1] + (i - (N - 1)).reshape(shape) * (x[N - 1] - x[N - 2])
            out = np.where((i < 0).reshape(shape), left, np.where((i >= N).reshape(shape), right, x[np.clip(i, 0, N - 1)]))
    else:
        raise ValueError("unknown signal extension mode %r" % (mode,))
    return out


def dwt_coeff_len(N, F, mode):
    """pywt.dwt_coeff_len (dictionaries/discrete.py:20-22)."""
    return (N + 1) // 2 if mode == "periodization" else (N + F - 1) // 2


def _dwt_ext(x, f_lo, f_hi, mode):
    """single-level DWT along axis 0 with signal extension ``mode`` (not periodization)."""
    N, F = x.shape[0], f_lo.shape[0]
    k = np.arange(dwt_coeff_len(N, F, mode))
    cA = np.zeros((k.shape[0],) + x.shape[1:], dtype=np.float64)
    cD = np.zeros_like(cA)
    for j in range(F):
        src = ext_sample(x, 2 * k + 1 - j, mode)
        cA += f_lo[j] * src
        cD += f_hi[j] * src
    return cA, cD


def _idwt_ext(cA, cD, rec_lo, rec_hi):
    """single-level inverse for every non-periodization mode: out[q] = sum_k cA[k] rec_lo[q + F - 2 - 2k] + cD[k] rec_hi[..],
    q = 0 .. 2 Nc - F + 1."""
    Nc, F = cA.shape[0], rec_lo.shape[0]
    out = np.zeros((2 * Nc - F + 2,) + cA.shape[1:], dtype=np.float64)
    q = np.arange(out.shape[0])
    shape = (-1,) + (1,) * (cA.ndim - 1)
    for k in range(Nc):
        t = q + F - 2 - 2 * k
        ok = (t >= 0) & (t < F)
        tt = np.clip(t, 0, F - 1)
        out += (np.where(ok, rec_lo[tt], 0.0).reshape(shape) * cA[k] + np.where(ok, rec_hi[tt], 0.0).reshape(shape) * cD[k])
    return out


def wavedec_mode(x, bank, mode, level=None):
    """pywt.wavedec(mode=...) along axis 0 -> [cA_L, cD_L, ..., cD_1] for any supported mode."""
    if mode == "periodization":
        return wavedec_per(x, bank, level)
    dec_lo, dec_hi = bank[0], bank[1]
    L = dwt_max_level(x.shape[0], dec_lo.shape[0]) if level is None else level
    a = np.asarray(x, dtype=np.float64)
    details = []
    for _ in range(L):
        a, d = _dwt_ext(a, dec_lo, dec_hi, mode)
        details.append(d)
    return [a] + details[::-1]


def waverec_mode(coeffs, bank, mode):
    """pywt.waverec(mode=...): the running approximation loses its last sample when it is one longer than the next detail
    band (pywt/_multilevel.py waverec); an odd-length signal therefore comes back one sample longer, as in PyWavelets."""
    if mode == "periodization":
        return waverec_per(coeffs, bank)
    rec_lo, rec_hi = bank[2], bank[3]
    a = coeffs[0]
    for d in coeffs[1:]:
        if a.shape[0] == d.shape[0] + 1:
            a = a[:-1]
        a = _idwt_ext(a, d, rec_lo, rec_hi)
    return a


class WaveletDict:
    """Analysis/synthesis pair on the REAL planar vector [Re|Im] as ONE signal (objectivefunction/priors/
    chi2.py:44-55), coefficients flattened like pywt.ravel_coeffs, optional ``append_signal``
    (dictionaries/undecimated.py:84-85,143-169; discrete.py:42-43,71-82)."""

    def __init__(self, name, kind="swt", level=None, norm=False, append_signal=False, fast_iswt=False,
                 mode="periodization"):
        self.bank = filter_bank(name)
        self.mode = mode  # decimated transform only (pywt signal-extension mode)
        self.kind, self.level, self.norm, self.append_signal = kind, level, norm, append_signal
        self.fast_iswt = fast_iswt  # closed-form iswt (same values; see iswt_fast)
        self.shapes = None
        self.N = None

    def decompose(self, x):
        self.N = x.shape[0]
        cs = swt(x, self.bank, self.level, self.norm) if self.kind == "swt" else \
            wavedec_mode(x, self.bank, self.mode, self.level)
        self.shapes = [c.shape[0] for c in cs]
        flat = np.concatenate(cs, axis=0)
        return np.concatenate([x, flat], axis=0) if self.append_signal else flat

    def reconstruct(self, c):
        sig = None
        if self.append_signal:
            sig, c = c[:self.N], c[self.N:]
        cs, o = [], 0
        for ln in self.shapes:
            cs.append(c[o:o + ln])
            o += ln
        if self.kind == "swt":
            out = iswt_fast(cs, self.bank, self.norm) if self.fast_iswt else iswt(cs, self.bank, self.norm)
        else:
            out = waverec_mode(cs, self.bank, self.mode)[:self.N]
        return sig + out if sig is not None else out


# ----------------------------------------------------------------------------------------------------------
# Chi2 / L1 / FISTA (objectivefunction/priors/chi2.py, l1.py, optimization/methods/fista.py)
# ----------------------------------------------------------------------------------------------------------
def soft_threshold(u, lam):
    """l1.py:30-32 on the real vector."""
    return np.sign(u) * np.maximum(np.abs(u) - lam, 0.0)


def chi2_gradient(op, data, z, wavelet=None, norm_factor=1.0):
    """chi2.py:43-56 ``calculate_gradient_fista``; also returns the model P_hat."""
    zs = wavelet.reconstruct(z) if wavelet is not None else z
    model = op.forward_normalized(real_to_complex(zs) * norm_factor)
    g = complex_to_real(op.backward(model - data))
    if wavelet is not None:
        g = wavelet.decompose(g)
    return g, model


def chi2_value(op, data, x, wavelet=None, norm_factor=1.0):
    """chi2.py:21-32 ``evaluate`` -> (0.5 sum w |d - model|^2, model)."""
    xs = wavelet.reconstruct(x) if wavelet is not None else x
    model = op.forward_normalized(real_to_complex(xs) * norm_factor)
    res = data - model
    w = op.w[:, None] if (op.w.ndim == 1 and res.ndim == 2) else op.w
    return 0.5 * np.sum(w * (res.real ** 2 + res.imag ** 2), axis=0), model


def l1_value(x):
    """l1.py:18-21 with epsilon = tiny(float32)."""
    return np.sum(np.sqrt(x * x + np.finfo(np.float32).tiny), axis=0)


def fista(op, data, x0, lam, noise, maxiter=None, wavelet=None, step=1.0, norm_factor=1.0):
    """fista.py:41-108.  ``lam`` / ``noise`` scalars or (P,) arrays with a COMMON iteration count; returns
    dict(x, cost, lam_final, model, iters).  ``step`` is the documented extension (1.0 = reference)."""
    x = np.array(x0, dtype=np.float64, copy=True)
    lam = np.array(lam, dtype=np.float64, copy=True)
    noise = np.asarray(1e-5 if noise is None else noise, dtype=np.float64)
    if maxiter is None:
        nz = np.where(noise != 0.0, noise, 1e-5)
        maxiter = int(np.floor(np.min(lam / nz)))
        noise = nz
    if np.all(noise >= lam):
        return dict(x=x, cost=np.zeros_like(lam), lam_final=lam, model=None, iters=0)
    t = 1.0
    z = x.copy()
    its = 0
    for _ in range(maxiter):
        xold = x
        g, _m = chi2_gradient(op, data, z, wavelet, norm_factor)
        x = soft_threshold(z - step * g, step * lam)
        t0 = t
        t = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t * t))
        z = x + ((t0 - 1.0) / t) * (x - xold)
        its += 1
        new = lam - noise
        if np.all(new > 0.0):
            lam = new
        else:
            break
    c2, model = chi2_value(op, data, x, wavelet, norm_factor)
    cost = c2 + lam * l1_value(x)
    return dict(x=x, cost=cost, lam_final=lam, model=model, iters=its)


# ----------------------------------------------------------------------------------------------------------
# synthetic sources (simulation/thinsource.py, thicksource.py, faradaysource.py)
# ----------------------------------------------------------------------------------------------------------
def thin_source(lambda2, s_nu, phi_gal, spectral_idx=0.0, dchi=0.0):
    """simulation/thinsource.py:21-29."""
    k, _ = spectral_factor(lambda2, spectral_idx)
    return s_nu * k * (np.cos(2.0 * phi_gal * lambda2) + 1j * np.sin(2.0 * phi_gal * lambda2 + dchi))


def thick_source(lambda2, s_nu, phi_fg, phi_center=0.0, spectral_idx=0.0):
    """simulation/thicksource.py:21-26."""
    k, _ = spectral_factor(lambda2, spectral_idx)
    return s_nu * k * np.exp(2j * lambda2 * phi_center) * np.sinc(phi_fg * lambda2 / np.pi)


def subtract_galactic_rm(data, lambda2, phi_gal):
    """base/dataset.py:377-381 (phi_gal scalar or (P,))."""
    ph = np.multiply.outer(lambda2, np.asarray(phi_gal, dtype=np.float64))
    return data * np.exp(-2j * ph)


# ----------------------------------------------------------------------------------------------------------
# after the loop: the README cube recipe (README.md:319-376) and the restore step
# (reconstruction/parameter.py:139-169, wrappers/reconstructors/csromer_reconstructor.py:59-84,205-213)
# ----------------------------------------------------------------------------------------------------------
def gaussian_beam(rmtf_fwhm, cellsize):
    """astropy.convolution.Gaussian1DKernel(stddev = round(sigma / cellsize)) as parameter.py:144-155 builds it.
    astropy is absent here: PARITY UNPINNED for this kernel (window = odd(8 sigma), unit-sum normalisation)."""
    sigma_pix = int(np.round(rmtf_fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0))) / cellsize))
    if sigma_pix < 1:
        return np.ones(1)
    size = int(np.ceil(8 * sigma_pix))
    size += 1 - size % 2
    x = np.arange(size) - size // 2
    k = np.exp(-0.5 * (x / float(sigma_pix)) ** 2) / (np.sqrt(2 * np.pi) * sigma_pix)
    return k / k.sum()


def convolve_beam(x, beam):
    """parameter.py:157-169: scipy.signal.convolve(mode="same", method="fft") on the real and imaginary parts"""
    from scipy import signal
    return signal.convolve(x.real, beam, mode="same", method="fft") + 1j * signal.convolve(x.imag, beam, mode="same",
                                                                                          method="fft")


def edge_noise(F_dirty, phi, max_faraday_depth, eta=1.0):
    """README.md:336-338"""
    edges = np.abs(phi) > max_faraday_depth / 1.5
    return eta * 0.5 * (np.std(F_dirty[edges].real) + np.std(F_dirty[edges].imag))


def peak_parabola(fd, cellsize):
    """csromer_reconstructor.py:59-78 -> (phi at the interpolated peak, interpolated peak)"""
    n = len(fd)
    i0 = int(np.argmax(np.abs(fd)))
    c0, m1, p1 = np.abs(fd[i0]), np.abs(fd[i0 - 1]), np.abs(fd[i0 + 1])
    pos = (p1 - m1) / (4 * c0 - 2 * m1 - 2 * p1)
    return (i0 + pos - n / 2) * cellsize, c0 - 0.25 * (m1 - p1) * pos


def cube_recipe(data, nu, sigma, spectral_idx=0.0, eta=1.0, wavelet="coif2", oversampling=8.0, step=1.0):
    """README.md:319-376 pixel by pixel: data (m, P) -> F (4, n, P) = dirty, model, restored, residual"""
    l2 = lambda2_from_nu(nu)
    sc = dataset_scalars(l2, sigma, spectral_idx)
    grid = phi_grid(l2, sc["w"], oversampling)
    phi, n, m = grid["phi"], grid["n"], len(l2)
    op = ExactDFT(l2, phi, sc["w"], sc["s"], sc["k"])
    beam = gaussian_beam(grid["rmtf_fwhm"], grid["cellsize"])
    P = data.shape[1]
    F = np.zeros((4, n, P), dtype=np.complex128)
    info = dict(noise=np.zeros(P), lambda_l1=np.zeros(P), rm_restored=np.zeros(P), rm_restored_interpolated=np.zeros(P),
                peak_restored=np.zeros(P))
    for p in range(P):
        d = data[:, p]
        dirty = op.backward(d)
        noise = edge_noise(dirty, phi, grid["max_faraday_depth"], eta)
        lam = np.sqrt(2 * m + np.sqrt(4 * m)) * noise
        wav = WaveletDict(wavelet, kind="swt") if wavelet else None
        x0 = complex_to_real(dirty)
        if wav is not None:
            x0 = wav.decompose(x0)
        r = fista(op, d, x0, lam, noise, None, wavelet=wav, step=step)
        xs = wav.reconstruct(r["x"]) if wav is not None else r["x"]
        model = real_to_complex(xs)
        f_res = op.backward(d - r["model"])
        restored = convolve_beam(model, beam) + f_res
        F[0, :, p], F[1, :, p], F[2, :, p], F[3, :, p] = dirty, model, restored, f_res
        info["noise"][p], info["lambda_l1"][p] = noise, lam
        info["rm_restored"][p] = phi[np.argmax(np.abs(restored))]
        info["rm_restored_interpolated"][p], info["peak_restored"][p] = peak_parabola(restored, grid["cellsize"])
    return F, info


# ----------------------------------------------------------------------------------------------------------
# Galactic RM foreground of an image (faraday_sky/faraday_sky.py:62-141).  astropy.wcs, astropy.coordinates and
# astropy_healpix are third-party packages absent here: PARITY UNPINNED against them; the chain below restates the
# published algorithms (FITS WCS papers I / II; IAU 1958 Galactic pole in J2000; Gorski et al. 2005 HEALPix) and is
# pinned by HEALPix known answers and the pix2ang round trip (tests/test_oracle_sky.py).
# ----------------------------------------------------------------------------------------------------------
def healpix_ang2pix_ring(nside, theta, phi):
    """HEALPix RING index of colatitude theta / longitude phi (radians)"""
    theta, phi = np.asarray(theta, dtype=np.float64), np.asarray(phi, dtype=np.float64)
    z = np.cos(theta)
    za = np.abs(z)
    tt = np.mod(phi, 2 * np.pi) / (0.5 * np.pi)
    ns = int(nside)
    npix = 12 * ns * ns
    # equatorial belt
    t1, t2 = ns * (0.5 + tt), ns * z * 0.75
    jp, jm = np.floor(t1 - t2).astype(np.int64), np.floor(t1 + t2).astype(np.int64)
    ir = ns + 1 + jp - jm
    kshift = 1 - (ir & 1)
    ip = np.mod((jp + jm - ns + kshift + 1) // 2, 4 * ns)
    eq = 2 * ns * (ns - 1) + (ir - 1) * 4 * ns + ip
    # polar caps
    tp = tt - np.floor(tt)
    tmp = ns * np.sqrt(3.0 * (1.0 - za))
    jp2, jm2 = np.floor(tp * tmp).astype(np.int64), np.floor((1.0 - tp) * tmp).astype(np.int64)
    ir2 = jp2 + jm2 + 1
    ip2 = np.mod(np.floor(tt * ir2).astype(np.int64), 4 * ir2)
    cap = np.where(z > 0, 2 * ir2 * (ir2 - 1) + ip2, npix - 2 * ir2 * (ir2 + 1) + ip2)
    return np.where(za <= 2.0 / 3.0, eq, cap)


def healpix_pix2ang_ring(nside, pix):
    """centre (theta, phi) of RING pixel `pix` -- the independent inverse used to pin ang2pix"""
    ns = int(nside)
    pix = np.asarray(pix, dtype=np.int64)
    npix, ncap = 12 * ns * ns, 2 * ns * (ns - 1)
    theta, phi = np.zeros(pix.shape), np.zeros(pix.shape)
    north = pix < ncap
    iring = ((1 + np.sqrt(1 + 2 * pix[north].astype(np.float64))) / 2).astype(np.int64)   # counted from the north pole
    iphi = pix[north] + 1 - 2 * iring * (iring - 1)
    theta[north] = np.arccos(1.0 - iring ** 2 / (3.0 * ns * ns))
    phi[north] = (iphi - 0.5) * np.pi / (2.0 * iring)
    eq = (pix >= ncap) & (pix < npix - ncap)
    ip = pix[eq] - ncap
    iring = ip // (4 * ns) + ns
    iphi = ip % (4 * ns) + 1
    fodd = 0.5 * (1 + ((iring + ns) & 1))
    theta[eq] = np.arccos((2 * ns - iring) * 2.0 / (3.0 * ns))
    phi[eq] = (iphi - fodd) * np.pi / (2.0 * ns)
    south = pix >= npix - ncap
    ip = npix - pix[south]
    iring = ((1 + np.sqrt(2 * ip.astype(np.float64) - 1)) / 2).astype(np.int64)            # counted from the south pole
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
    theta[south] = np.arccos(-1.0 + iring ** 2 / (3.0 * ns * ns))
    phi[south] = (iphi - 0.5) * np.pi / (2.0 * iring)
    return theta, phi


def icrs_to_galactic(ra, dec):
    """(l, b) in radians from right ascension / declination (J2000 pole 192.85948, 27.12825 deg; l_NCP = 122.93192 deg)"""
    ra_ngp, dec_ngp, l_ncp = np.radians(192.85948), np.radians(27.12825), np.radians(122.93192)
    b = np.arcsin(np.clip(np.sin(dec) * np.sin(dec_ngp) + np.cos(dec) * np.cos(dec_ngp) * np.cos(ra - ra_ngp), -1, 1))
    l = l_ncp - np.arctan2(np.cos(dec) * np.sin(ra - ra_ngp),
                           np.sin(dec) * np.cos(dec_ngp) - np.cos(dec) * np.sin(dec_ngp) * np.cos(ra - ra_ngp))
    return np.mod(l, 2 * np.pi), b


def wcs_pixel_to_sky(x, y, crval, crpix, cd, proj="TAN", lonpole=180.0):
    """0-based pixel (x, y) -> (ra, dec) radians for a zenithal TAN / SIN WCS (FITS WCS paper II, eqs. 2, 54-59)"""
    px, py = np.asarray(x, dtype=np.float64) + 1.0 - crpix[0], np.asarray(y, dtype=np.float64) + 1.0 - crpix[1]
    xi = np.radians(cd[0][0] * px + cd[0][1] * py)
    eta = np.radians(cd[1][0] * px + cd[1][1] * py)
    rr = np.hypot(xi, eta)
    phi = np.where(rr > 0, np.arctan2(xi, -eta), 0.0)
    theta = np.arccos(np.minimum(rr, 1.0)) if proj == "SIN" else np.arctan2(1.0, rr)
    a0, d0, dp = np.radians(crval[0]), np.radians(crval[1]), phi - np.radians(lonpole)
    dec = np.arcsin(np.clip(np.sin(theta) * np.sin(d0) + np.cos(theta) * np.cos(d0) * np.cos(dp), -1, 1))
    ra = a0 + np.arctan2(-np.cos(theta) * np.sin(dp), np.sin(theta) * np.cos(d0) - np.cos(theta) * np.sin(d0) * np.cos(dp))
    return ra, dec


def galactic_rm_image(sky_mean, sky_std, nside, header, swap_axes=False):
    """faraday_sky.py:99-141 (nearest HEALPix pixel, RING): -> (rm_mean [NAXIS2, NAXIS1], rm_std)"""
    nx, ny = int(header["NAXIS1"]), int(header["NAXIS2"])
    cd = [[header.get("CD1_1", header.get("CDELT1", 1.0)), header.get("CD1_2", 0.0)],
          [header.get("CD2_1", 0.0), header.get("CD2_2", header.get("CDELT2", 1.0))]]
    proj = "SIN" if str(header.get("CTYPE1", "RA---TAN")).endswith("SIN") else "TAN"
    cc, rr = np.meshgrid(np.arange(nx), np.arange(ny))
    x, y = (rr, cc) if swap_axes else (cc, rr)
    ra, dec = wcs_pixel_to_sky(x, y, (header["CRVAL1"], header["CRVAL2"]), (header["CRPIX1"], header["CRPIX2"]), cd, proj,
                               header.get("LONPOLE", 180.0))
    l, b = icrs_to_galactic(ra, dec)
    pix = healpix_ang2pix_ring(nside, 0.5 * np.pi - b, l)
    return np.asarray(sky_mean)[pix], np.asarray(sky_std)[pix]
