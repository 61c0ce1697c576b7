"""Host-side mirror of the reference API (Dataset / Parameter / simulation / OFunction / L1 / filters) against the
golden fixtures the reference produced, plus the reference's own (four) unit tests."""
import numpy as np
import pytest
from conftest import GOLDEN_CASES, golden

import csromer_b200 as cs
from csromer_b200.dictionaries import filters
from csromer_b200.optimization.methods.fista import _decayed_lambda
from oracle import restatement as ob


class TestEmptyDataset:
    """the reference's tests/unit/base/test_dataset.py:7-17, verbatim semantics"""
    dataset = cs.Dataset()

    def test_empty_dataset(self):
        assert self.dataset is not None

    def test_empty_dataset_k(self):
        assert self.dataset.k is None

    def test_empty_dataset_nu(self):
        assert self.dataset.nu is None

    def test_empty_dataset_lambda2(self):
        assert self.dataset.lambda2 is None


def make_source(name, g):
    nu = g["nu"]
    thin = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    thin.simulate()
    src = thin
    if name.startswith("mixed"):
        thick = cs.FaradayThickSource(nu=nu, s_nu=0.0035, phi_fg=140, phi_center=200, spectral_idx=1.0)
        thick.simulate()
        src = thin + thick
    if "noise" in name and "nonoise" not in name:
        seed = int(name.split("noise")[1].split("_")[0])
        src.apply_noise(0.2 * 0.0035, random_state=np.random.RandomState(seed))
    return src


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_dataset_matches_reference(name):
    g = golden(name)
    src = make_source(name, g)
    assert src.m == len(g["lambda2"])
    np.testing.assert_array_equal(src.lambda2, g["lambda2"])
    np.testing.assert_allclose(src.data, g["data"], rtol=1e-13, atol=1e-18)  # same RNG stream as the reference
    np.testing.assert_allclose(src.w, g["w"], rtol=1e-13)
    np.testing.assert_allclose(src.sigma, g["sigma"], rtol=1e-13)
    np.testing.assert_allclose(src.s, g["s"], rtol=1e-13)
    np.testing.assert_allclose(src.k, g["k"], rtol=1e-13)
    assert src.l2_ref == float(g["l2_ref"])
    np.testing.assert_allclose(src.spectral_idx, g["spectral_idx"], rtol=1e-14)
    if np.isnan(g["theo_noise"]):
        assert src.theo_noise is None
    else:
        np.testing.assert_allclose(src.theo_noise, g["theo_noise"], rtol=1e-13)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_parameter_matches_reference(name, capsys):
    g = golden(name)
    src = make_source(name, g)
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8)
    assert "FWHM of the main peak of the RMTF" in capsys.readouterr().out
    assert par.n == int(g["n"]) and len(par.phi) == int(g["n"])
    np.testing.assert_array_equal(par.phi, g["phi"])
    for key in ("cellsize", "rmtf_fwhm", "max_recovered_width", "max_faraday_depth"):
        np.testing.assert_allclose(getattr(par, key), g[key], rtol=1e-14)
    assert par.data.dtype == np.complex64 and par.data.shape == (par.n,)
    # packing round trip and the n-overwrite quirk (parameter.py:46-50)
    par.data = (np.arange(par.n) + 1j * np.arange(par.n)).astype(np.complex64)
    par.complex_data_to_real()
    assert par.n == 2 * len(par.phi) and par.data.dtype == np.float32
    with pytest.raises(TypeError):
        par.complex_data_to_real()
    par.real_data_to_complex()
    assert par.n == len(par.phi)
    with pytest.raises(ValueError):
        par.real_data_to_complex()


def test_dataset_setter_semantics():
    nu = np.linspace(1.0e9, 2.0e9, 16)
    ds = cs.Dataset(nu=nu, data=np.ones(16, dtype=np.complex64), sigma=np.full(16, 0.5))
    assert np.all(np.diff(ds.lambda2) > 0) and np.all(np.diff(ds.nu) < 0)
    np.testing.assert_allclose(ds.w, 4.0)
    np.testing.assert_allclose(ds.k, 64.0)
    np.testing.assert_allclose(ds.theo_noise, 1 / 8.0)
    # in-place flagging bypasses the setter: k stays stale, exactly like the reference's flaggers leave it
    ds.w[:4] = 0.0
    np.testing.assert_allclose(ds.k, 64.0)
    # assigning the array refreshes everything
    ds.w = ds.w.copy()
    np.testing.assert_allclose(ds.k, 48.0)
    assert ds.sigma[0] == 0.0 and ds.sigma[-1] == 0.5
    ds.model_data = np.full(16, 0.25, dtype=np.complex64)
    np.testing.assert_allclose(ds.residual, 0.75)
    with pytest.raises(SystemExit):
        ds.model_data = np.zeros(3)
    ds.spectral_idx = 0.7
    np.testing.assert_allclose(ds.s, (ds.nu / ds.nu_0) ** 0.7)
    np.testing.assert_allclose(ds.k, np.sum(ds.w / ds.s))
    desc = cs.Dataset(lambda2=np.linspace(0.09, 0.02, 8))
    assert np.all(np.diff(desc.lambda2) > 0)
    np.testing.assert_allclose(desc.calculate_l2ref(), np.mean(desc.lambda2))


def test_batched_dataset():
    nu = np.linspace(1.0e9, 2.0e9, 32)
    rng = np.random.RandomState(0)
    sig = rng.uniform(0.5, 2.0, size=(32, 3, 4))
    ds = cs.Dataset(nu=nu, sigma=sig, data=np.ones((32, 3, 4), dtype=np.complex64), spectral_idx=0.3)
    assert ds.m == 32 and ds.n_pixels == 12 and ds.k.shape == (3, 4) and ds.theo_noise.shape == (3, 4)
    np.testing.assert_allclose(ds.k, np.sum(ds.w / ds.s[:, None, None], axis=0))
    phi_gal = rng.normal(0, 30, size=(3, 4))
    before = ds.data.copy()
    ds.subtract_galacticrm(phi_gal)
    np.testing.assert_allclose(ds.data, ob.subtract_galactic_rm(before, ds.lambda2, phi_gal), rtol=1e-13)


def test_batched_sources_match_oracle():
    nu = np.linspace(1.008e9, 2.031e9, 64)
    phi = np.array([-200.0, 35.5, 410.0])
    amp = np.array([0.0035, 0.002, 0.004])
    thin = cs.FaradayThinSource(nu=nu, s_nu=amp, phi_gal=phi, spectral_idx=1.0)
    thin.simulate()
    assert thin.data.shape == (64, 3)
    for p in range(3):
        np.testing.assert_allclose(thin.data[:, p], ob.thin_source(thin.lambda2, amp[p], phi[p], 1.0), rtol=1e-13)
    thick = cs.FaradayThickSource(nu=nu, s_nu=0.0035, phi_fg=140, phi_center=200, spectral_idx=1.0)
    thick.simulate()
    np.testing.assert_allclose(thick.data, ob.thick_source(thick.lambda2, 0.0035, 140, 200, 1.0), rtol=1e-13)


def test_remove_channels_deletes():
    nu = np.linspace(1.008e9, 2.031e9, 500)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    src.simulate()
    src.remove_channels(0.2, random_state=np.random.RandomState(3))
    assert src.m == len(src.lambda2) == len(src.data)
    assert abs(src.m / 500.0 - 0.8) < 0.01


def test_ofunction_and_l1():
    l1 = cs.L1(reg=0.5)
    x = np.array([-2.0, -0.3, 0.0, 0.4, 3.0])
    np.testing.assert_allclose(l1.calculate_prox(x), [-1.5, 0, 0, 0, 2.5])
    np.testing.assert_allclose(l1.evaluate(x), np.abs(x).sum(), rtol=1e-12)
    f = cs.OFunction([l1])
    assert f.getLambda() == 0.5
    f.setLambda(reg=0.25)
    assert l1.reg == 0.25
    np.testing.assert_allclose(f.evaluate(x), 0.25 * np.abs(x).sum(), rtol=1e-12)
    np.testing.assert_allclose(f.calc_prox(x), ob.soft_threshold(x, 0.25))
    assert cs.Chi2(dft_obj=None).reg == 1.0 and cs.L1().norm_factor == 1.0


def test_decayed_lambda_matches_loop():
    g = golden("thin_noise0_jvla1000_k100")
    assert _decayed_lambda(float(g["lam0"]), float(g["noise"]), 100) == float(g["lam_final"])
    # break branch: lambda - noise <= 0 stops the decay (fista.py:100-106)
    assert _decayed_lambda(1.0, 0.3, 100) == pytest.approx(0.1)
    arr = _decayed_lambda(np.array([1.0, 1.0]), np.array([0.3, 0.1]), 5)
    np.testing.assert_allclose(arr, [0.1, 0.5])


def test_filter_banks_agree_with_oracle_tables():
    for name in ["db1", "haar", "db2", "sym2", "db3", "db4", "coif1", "coif2", "sym4", "bior2.2", "IUWT"]:
        for mine, theirs in zip(filters.filter_bank(name), ob.filter_bank(name)):
            np.testing.assert_allclose(mine, theirs, rtol=0, atol=0)
    with pytest.raises(NotImplementedError):
        filters.filter_bank("nope")
    assert filters.swt_max_level(15936) == 6 and filters.swt_max_level(31936) == 6
    assert filters.dwt_max_level(15936, 12) == ob.dwt_max_level(15936, 12)


def test_wavelet_constructor_errors():
    with pytest.raises(TypeError):
        cs.UndecimatedWavelet(wavelet_name=3)
    with pytest.raises(TypeError):
        cs.DiscreteWavelet(wavelet_name="db2", mode=None)
    with pytest.raises(ValueError):
        cs.DiscreteWavelet(wavelet_name="db2", mode="mirror")
    for mode in ("periodization", "symmetric", "zero", "constant", "reflect", "periodic", "smooth", "antisymmetric", "antireflect"):
        w = cs.DiscreteWavelet(wavelet_name="db2", mode=mode)
        assert w.calculate_ncoeffs(np.zeros(9)) == (5 if mode == "periodization" else 6)  # pywt.dwt_coeff_len(9, 4, mode)
    with pytest.raises(NotImplementedError):
        cs.UndecimatedWavelet(wavelet_name="not-a-wavelet")
    w = cs.UndecimatedWavelet(wavelet_name="IUWT")
    assert w.trim_approx is True and w.norm is False and len(w.wavelet[0]) == 6


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from csromer_b200._lib import CsrError
    nu = np.linspace(1.008e9, 2.031e9, 32)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    src.simulate()
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, verbose=False)
    with pytest.raises(CsrError):
        cs.NDFT1D(dataset=src, parameter=par).backward(src.data)


def test_dataset_stack_presents_ragged_datasets_as_one_batch():
    """host logic of the per-line-of-sight sampling path (PerPixelDFT1D): zero padding, fan-out of model_data"""
    from csromer_b200.transformers.dfts import DatasetStack
    rng = np.random.RandomState(0)
    nu = np.linspace(1.0e9, 2.0e9, 40)
    full = cs.Dataset(nu=nu, sigma=np.ones(40))
    members = []
    for p, keep in enumerate((40, 31, 25)):
        idx = np.sort(rng.choice(40, size=keep, replace=False))
        ds = cs.Dataset(lambda2=np.asarray(full.lambda2)[idx], sigma=0.1 * (1 + rng.uniform(size=keep)),
                        data=(rng.normal(size=keep) + 1j * rng.normal(size=keep)).astype(np.complex64),
                        spectral_idx=0.5, l2_ref=0.01 * p)
        members.append(ds)
    st = DatasetStack(members)
    assert (st.m, st.P) == (40, 3) and list(st.m_count) == [40, 31, 25]
    assert st.data.shape == (40, 3) and np.all(st.data[31:, 1] == 0) and np.all(st.w[25:, 2] == 0)
    np.testing.assert_array_equal(st.data[:25, 2], members[2].data)
    np.testing.assert_allclose(st.lambda2_minus_ref[:31, 1], np.asarray(members[1].lambda2) - 0.01)
    np.testing.assert_allclose(st.k, [float(d.k) for d in members])
    np.testing.assert_allclose(st.theo_noise, [d.theo_noise for d in members])
    assert np.all(st.s[31:, 1] == 1.0)
    model = (rng.normal(size=(40, 3)) + 1j * rng.normal(size=(40, 3)))
    st.model_data = model
    for p, ds in enumerate(members):
        np.testing.assert_allclose(ds.model_data, model[:ds.m, p].astype(np.complex64))
        np.testing.assert_allclose(ds.residual, ds.data - ds.model_data)
    np.testing.assert_allclose(st.residual[:25, 2], members[2].residual)
    with pytest.raises(ValueError):
        DatasetStack([])
    with pytest.raises(ValueError):
        DatasetStack([cs.Dataset(nu=nu, sigma=np.ones(40), data=np.zeros((40, 2), np.complex64))])


def test_make_mask_helpers_follow_the_reference():
    """utils/utilities.py:35-65 of the reference: np.where tuples of kept / masked pixels"""
    from csromer_b200.utils import make_mask, make_mask_faraday
    rng = np.random.RandomState(3)
    I, P = rng.uniform(0, 1, (5, 6)), rng.uniform(0, 1, (5, 6))
    Q, U = rng.normal(size=(7, 5, 6)), rng.normal(size=(7, 5, 6))
    Q[3, 1, 2] = np.nan
    U[0, 4, 5] = np.nan
    alpha = rng.normal(size=(5, 6))
    alpha[2, 2] = np.nan
    kept, masked = make_mask(I, 0.4)
    assert set(zip(*kept)) == set(zip(*np.where(I >= 0.4))) and len(kept[0]) + len(masked[0]) == I.size
    kept, masked = make_mask_faraday(I, P, Q, U, None, 0.3, 0.2)
    expect = (I >= 0.3) & (P >= 0.2) & ~np.isnan(Q).any(axis=0) & ~np.isnan(U).any(axis=0)
    assert set(zip(*kept)) == set(zip(*np.where(expect))) and (1, 2) not in set(zip(*kept)) and (4, 5) in set(zip(*masked))
    kept2, _ = make_mask_faraday(I, P, Q, U, alpha, 0.3, 0.2)
    assert set(zip(*kept2)) == set(zip(*np.where(expect & ~np.isnan(alpha))))


def test_parameter_remembers_which_complex_tensor_its_planar_copy_came_from():
    """Parameter.complex_source(): the complex torch tensor `data` was made from by complex_data_to_real, as long as neither has
    been modified (version counters) or replaced -- what lets FISTA reuse the planar device copy NDFT1D.backward kept."""
    import torch
    par = cs.Parameter()
    z = torch.complex(torch.arange(6.0), -torch.arange(6.0))
    par.data = z
    assert par.complex_source() is None
    par.complex_data_to_real()
    assert par.data.shape == (12,) and par.complex_source() is z
    par.data[3] = 7.0                      # in place: the version counter of the planar copy moves
    assert par.complex_source() is None
    par.data = z
    par.complex_data_to_real()
    z.mul_(2.0)                            # ... or the source's
    assert par.complex_source() is None
    par.data = z
    par.complex_data_to_real()
    par.data = par.data.clone()            # replaced
    assert par.complex_source() is None
    import copy, pickle
    par.data = z
    par.complex_data_to_real()
    assert copy.deepcopy(par).data.shape == (12,)
    assert pickle.loads(pickle.dumps(par)).complex_source() is None
    npar = cs.Parameter()
    npar.data = np.arange(4) + 1j * np.arange(4)
    npar.complex_data_to_real()            # numpy arrays carry no version: never "the same"
    assert npar.complex_source() is None
