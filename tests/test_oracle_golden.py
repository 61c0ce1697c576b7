"""Pin Oracle B (oracle/restatement.py, float64 NumPy) to Oracle A (the reference's own classes, frozen in
tests/golden/ by oracle/make_golden.py).  Tolerance 2e-7 relative to the peak: the reference rounds every
transform to complex64 (ndft.py:19,25,34,42), the restatement does not."""
import numpy as np
import pytest
from conftest import GOLDEN_CASES, SMALL_CASES, golden

from oracle import restatement as ob

TOL = 2e-7


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_scalars_and_grid(name):
    g = golden(name)
    sc = ob.dataset_scalars(g["lambda2"], g["sigma"], float(g["spectral_idx"]))
    np.testing.assert_allclose(sc["w"], g["w"], rtol=1e-13)
    np.testing.assert_allclose(sc["s"], g["s"], rtol=1e-13)
    np.testing.assert_allclose(sc["k"], g["k"], rtol=1e-13)
    if np.isnan(g["theo_noise"]):
        assert sc["theo_noise"] is None
    else:
        np.testing.assert_allclose(sc["theo_noise"], g["theo_noise"], rtol=1e-13)
    np.testing.assert_array_equal(ob.lambda2_from_nu(g["nu"]), g["lambda2"])
    pg = ob.phi_grid(g["lambda2"], g["w"], 8.0)
    assert pg["n"] == int(g["n"])
    np.testing.assert_array_equal(pg["phi"], g["phi"])
    for key in ("cellsize", "rmtf_fwhm", "max_recovered_width", "max_faraday_depth"):
        np.testing.assert_allclose(pg[key], g[key], rtol=1e-14)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_transforms(name):
    g = golden(name)
    op = ob.ExactDFT(g["lambda2"], g["phi"], g["w"], g["s"], float(g["k"]), float(g["l2_ref"]))
    assert rel(op.backward(g["data"]), g["F_dirty"]) < TOL
    assert rel(op.rmtf(), g["rmtf"]) < TOL
    assert rel(op.forward(g["x_test"]), g["fwd_x_test"]) < TOL
    assert rel(op.forward_normalized(g["x_test"]), g["fwdn_x_test"]) < TOL


def test_reference_known_answers():
    """numbers quoted in SURVEY.md section 8c, produced by the reference itself"""
    g = golden("thin_nonoise_jvla1000")
    assert int(g["n"]) == 7968 and len(g["lambda2"]) == 1000
    np.testing.assert_allclose(g["k"], 1040.6553522078002, rtol=1e-12)
    np.testing.assert_allclose(g["cellsize"], 6.514783041698439, rtol=1e-12)
    j = np.argmax(np.abs(g["F_dirty"]))
    np.testing.assert_allclose(g["phi"][j], -201.9582742926516, rtol=1e-12)
    np.testing.assert_allclose(np.abs(g["F_dirty"]).max(), 0.003354779466042806, rtol=1e-6)
    np.testing.assert_allclose(np.abs(g["rmtf"]).max(), 1.0, atol=1e-6)


@pytest.mark.parametrize("name", SMALL_CASES + ["mixed_noise1_jvla1000_k30"])
def test_fista(name):
    g = golden(name)
    op = ob.ExactDFT(g["lambda2"], g["phi"], g["w"], g["s"], float(g["k"]), float(g["l2_ref"]))
    x0 = ob.complex_to_real(g["F_dirty"].astype(np.complex128))
    r = ob.fista(op, g["data"], x0, float(g["lam0"]), float(g["noise"]), int(g["maxiter"]))
    assert np.isfinite(r["x"]).all()
    assert rel(r["x"], g["fista_x"]) < TOL
    assert rel(r["model"], g["model_data"]) < TOL
    np.testing.assert_allclose(r["lam_final"], g["lam_final"], rtol=1e-12)
    np.testing.assert_allclose(r["cost"], g["fista_obj"], rtol=1e-6)
    assert np.count_nonzero(r["x"]) == np.count_nonzero(g["fista_x"])


def test_fista_batch_equals_single():
    """the batched restatement (pixel-last arrays) reproduces the single-LOS runs column by column"""
    g = golden("mixed_noise1_jvla200_k60")
    rng = np.random.RandomState(0)
    scale = rng.uniform(0.5, 2.0, size=3)
    data = g["data"][:, None] * scale[None, :]
    op = ob.ExactDFT(g["lambda2"], g["phi"], g["w"], g["s"], float(g["k"]))
    x0 = ob.complex_to_real(op.backward(data))
    lam = float(g["lam0"]) * scale
    nz = float(g["noise"]) * scale
    rb = ob.fista(op, data, x0, lam, nz, 20)
    for p in range(3):
        rs = ob.fista(op, data[:, p], x0[:, p], lam[p], nz[p], 20)
        assert rel(rb["x"][:, p], rs["x"]) < 1e-12
