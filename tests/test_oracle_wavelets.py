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
