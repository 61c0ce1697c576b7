"""The wavelet restatement (PyWavelets semantics; PyWavelets is absent, parity unpinned) is held to the
identities SURVEY.md section 7.1 lists: pywt-documentation known answers, perfect reconstruction, Parseval for
orthogonal banks, shift equivariance of the undecimated transform, tight-frame energy with norm=True."""
import numpy as np
import pytest

from oracle import restatement as ob


def test_pywt_doc_known_answers():
    s2 = np.sqrt(2.0)
    cA, cD = ob._dwt_per(np.arange(1, 7.0), *ob.filter_bank("db1")[:2])
    np.testing.assert_allclose(cA, [3 / s2, 7 / s2, 11 / s2], rtol=1e-14)
    np.testing.assert_allclose(cD, [-1 / s2] * 3, rtol=1e-14)
    c = ob.wavedec_per(np.arange(1, 9.0), ob.filter_bank("db1"), 2)
    np.testing.assert_allclose(c[0], [5, 13], rtol=1e-13)
    np.testing.assert_allclose(c[1], [-2, -2], rtol=1e-13)
    np.testing.assert_allclose(c[2], [-1 / s2] * 4, rtol=1e-13)
    a = ob.swt(np.arange(1, 9.0), ob.filter_bank("db1"), 1)[0]
    np.testing.assert_allclose(a, np.array([3, 5, 7, 9, 11, 13, 15, 9]) / s2, rtol=1e-13)


@pytest.mark.parametrize("name", ["db1", "db2", "db3", "db4", "sym4", "coif1", "coif2"])
def test_orthonormal_banks(name):
    dec_lo, dec_hi, rec_lo, rec_hi = ob.filter_bank(name)
    np.testing.assert_allclose(np.sum(dec_lo), np.sqrt(2), rtol=1e-9)
    np.testing.assert_allclose(np.sum(dec_lo ** 2), 1.0, rtol=1e-9)
    np.testing.assert_allclose(np.sum(dec_hi), 0.0, atol=1e-9)
    for shift in range(2, len(dec_lo), 2):
        assert abs(np.dot(dec_lo[shift:], dec_lo[:-shift])) < 1e-9


@pytest.mark.parametrize("name,kind,N,level,norm,app", [
    ("db1", "swt", 64, None, False, False), ("coif2", "swt", 3136, None, False, False),
    ("sym2", "swt", 3136, 3, False, True), ("db4", "swt", 3136, None, True, False),
    ("bior2.2", "swt", 1024, 4, False, False), ("db2", "dwt", 3136, None, False, False),
    ("coif2", "dwt", 15936, None, False, False), ("db1", "dwt", 250, None, False, True),
])
def test_perfect_reconstruction(name, kind, N, level, norm, app):
    rng = np.random.RandomState(1)
    x = rng.normal(size=(N, 2))
    w = ob.WaveletDict(name, kind=kind, level=level, norm=norm, append_signal=app)
    c = w.decompose(x)
    r = w.reconstruct(c)
    np.testing.assert_allclose(r, 2 * x if app else x, atol=5e-10)
    if kind == "swt":
        L = ob.swt_max_level(N) if level is None else level
        assert c.shape[0] == (L + 1) * N + (N if app else 0)


def test_dwt_odd_levels_coefficient_count():
    x = np.random.RandomState(0).normal(size=15936)
    c = ob.WaveletDict("coif2", kind="dwt").decompose(x)
    assert c.shape[0] == 15939  # 249 -> 125 -> 63 -> 32 ...


def test_parseval_and_tight_frame():
    x = np.random.RandomState(3).normal(size=4096)
    c = ob.WaveletDict("db4", kind="dwt").decompose(x)
    np.testing.assert_allclose(np.sum(c ** 2), np.sum(x ** 2), rtol=1e-9)
    c = ob.WaveletDict("db4", kind="swt", norm=True).decompose(x)
    np.testing.assert_allclose(np.sum(c ** 2), np.sum(x ** 2), rtol=1e-9)


def test_swt_shift_equivariance():
    x = np.random.RandomState(4).normal(size=512)
    bank = ob.filter_bank("coif1")
    a = ob.swt(x, bank, 4)
    b = ob.swt(np.roll(x, 5), bank, 4)
    for ca, cb in zip(a, b):
        np.testing.assert_allclose(np.roll(ca, 5), cb, atol=1e-12)


def test_against_pywt_fixtures():
    """the pin: fixtures produced by PyWavelets itself (oracle/make_golden_wavelets.py), when somebody could make them"""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wavelets_pywt.npz")
    if not os.path.exists(path):
        pytest.skip("no pywt-produced fixtures (PyWavelets is not installed here): wavelet parity stays unpinned")
    g = np.load(path)
    for key in [k[:-5] for k in g.files if k.endswith("_meta")]:
        kind, name, N, lev, norm = g[key + "_meta"]
        bank = ob.filter_bank(str(name))
        x = g[key + "_x"]
        if kind == "swt":
            cs = ob.swt(x, bank, int(lev), bool(int(norm)))
            rec = ob.iswt(cs, bank, bool(int(norm)))
        else:
            cs = ob.wavedec_per(x, bank, int(lev))
            rec = ob.waverec_per(cs, bank)
        np.testing.assert_array_equal([len(c) for c in cs], g[key + "_sizes"])
        np.testing.assert_allclose(np.concatenate(cs), g[key + "_flat"], rtol=0, atol=1e-12 * np.abs(g[key + "_flat"]).max())
        np.testing.assert_allclose(rec[:len(g[key + "_rec"])], g[key + "_rec"], rtol=0, atol=1e-12 * np.abs(x).max())


def test_closed_form_iswt_equals_the_literal_one():
    """iswt_fast (one a-trous gather per level) == iswt (pywt's double loop over phases), every bank, odd level counts"""
    rng = np.random.RandomState(0)
    for name in ("db1", "coif2", "sym4", "bior2.2", "IUWT", "db3"):
        for N, lev in ((256, None), (480, 3), (1536, None)):
            bank = ob.filter_bank(name)
            L = ob.swt_max_level(N) if lev is None else lev
            cs = [rng.normal(size=(N, 3)) for _ in range(L + 1)]
            for norm in (False, True):
                np.testing.assert_allclose(ob.iswt_fast(cs, bank, norm), ob.iswt(cs, bank, norm), rtol=0, atol=1e-13)


# ---- decimated transform with signal-extension modes --------------------------------------------------------
def test_dwt_symmetric_matches_the_pywavelets_documentation_value():
    """PyWavelets' documentation (pywt.dwt, "Single level dwt"): ``pywt.dwt([3, 7, 1, 1, -2, 5, 4, 6], 'db2')`` with the
    default mode 'symmetric' prints these ten numbers -- a known answer that pins the filter orientation, the odd-position
    sampling and the band length (N + F - 1) // 2 of the restatement."""
    bank = ob.filter_bank("db2")
    cA, cD = ob._dwt_ext(np.array([3, 7, 1, 1, -2, 5, 4, 6.0]), bank[0], bank[1], "symmetric")
    np.testing.assert_allclose(cA, [5.65685425, 7.39923721, 0.22414387, 3.33677403, 7.77817459], atol=5e-9)
    np.testing.assert_allclose(cD, [-2.44948974, -1.60368225, -4.44140056, -0.41361256, 1.22474487], atol=5e-9)


def test_signal_extensions_match_numpy_pad():
    """pywt.pad documents the correspondence of its modes with numpy.pad (zero = constant, constant = edge, periodic = wrap,
    antireflect = reflect with reflect_type='odd'); extensions longer than the signal keep following numpy's repetition."""
    rng = np.random.default_rng(0)
    kw = {"zero": dict(mode="constant"), "constant": dict(mode="edge"), "symmetric": dict(mode="symmetric"),
          "reflect": dict(mode="reflect"), "periodic": dict(mode="wrap"), "antireflect": dict(mode="reflect", reflect_type="odd")}
    for N in (1, 2, 5, 8):
        x = rng.normal(size=N)
        for mode, k in kw.items():
            if N == 1 and mode in ("reflect", "antireflect"):
                continue
            for pad in (1, 3, max(N - 1, 1), 2 * N + 3):
                np.testing.assert_allclose(ob.ext_sample(x, np.arange(-pad, N + pad), mode), np.pad(x, pad, **k), atol=1e-12)
    x = np.array([1.0, 2.0, 4.0])
    np.testing.assert_allclose(ob.ext_sample(x, np.arange(-2, 5), "smooth"), [-1, 0, 1, 2, 4, 6, 8])
    np.testing.assert_allclose(ob.ext_sample(x, np.arange(-2, 5), "antisymmetric"), [-2, -1, 1, 2, 4, -4, -2])


@pytest.mark.parametrize("mode", ob.EXT_MODES)
def test_extension_modes_reconstruct_perfectly(mode):
    rng = np.random.default_rng(1)
    for name in ("db1", "db2", "coif2", "sym4"):
        bank = ob.filter_bank(name)
        F = bank[0].shape[0]
        for N in (16, 17, 100, 3136):
            x = rng.normal(size=(N, 2))
            cs_ = ob.wavedec_mode(x, bank, mode)
            n = N
            for band in cs_[:0:-1]:      # cD_1 .. cD_L
                n = ob.dwt_coeff_len(n, F, mode)
                assert band.shape[0] == n
            assert cs_[0].shape[0] == n
            r = ob.waverec_mode(cs_, bank, mode)
            assert r.shape[0] in (N, N + 1) and (N % 2 or r.shape[0] == N)
            assert np.abs(r[:N] - x).max() < 5e-8  # filter tables hold ~1e-10; "smooth" extrapolates to large edge values
    d = ob.WaveletDict("coif2", kind="dwt", mode=mode, append_signal=True)
    x = rng.normal(size=(200, 3))
    assert np.abs(d.reconstruct(d.decompose(x)) - 2 * x).max() < 1e-9


def test_multilevel_symmetric_matches_the_pywavelets_documentation_values():
    """two more documented PyWavelets answers (default mode 'symmetric'): `pywt.wavedec([1..8], 'db1', level=2)` prints
    cA2 = [5, 13], cD2 = [-2, -2], cD1 = [-0.70710678] * 4, and `pywt.dwt([1..6], 'db1')` prints
    ([2.12132034, 4.94974747, 7.77817459], [-0.70710678] * 3)"""
    bank = ob.filter_bank("db1")
    cA2, cD2, cD1 = ob.wavedec_mode(np.arange(1.0, 9.0), bank, "symmetric", level=2)
    np.testing.assert_allclose(cA2, [5.0, 13.0], atol=1e-12)
    np.testing.assert_allclose(cD2, [-2.0, -2.0], atol=1e-12)
    np.testing.assert_allclose(cD1, [-0.70710678] * 4, atol=5e-9)
    cA, cD = ob.wavedec_mode(np.arange(1.0, 7.0), bank, "symmetric", level=1)
    np.testing.assert_allclose(cA, [2.12132034, 4.94974747, 7.77817459], atol=5e-9)
    np.testing.assert_allclose(cD, [-0.70710678] * 3, atol=5e-9)
    np.testing.assert_allclose(ob.waverec_mode([cA2, cD2, cD1], bank, "symmetric"), np.arange(1.0, 9.0), atol=1e-12)
