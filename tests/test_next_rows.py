"""Rows next to the hot path: flaggers, calculate_noise, filter_cubes, cube I/O layouts.

CPU part: this package's host mirrors against tests/golden/next_rows.npz, frozen from the reference's own classes by
oracle/make_golden_next.py; the FITS codec against the standard's layout and a round trip.
GPU part: the cube transposes and channel statistics of csrc/cubeio.cu against NumPy on the same inputs (bit exact for
the transposes, 1e-6 relative for the fp64-accumulated statistics)."""
import os

import numpy as np
import pytest
from conftest import golden

from oracle import restatement as ob

import csromer_b200 as cs
from csromer_b200.io import fits_lite
from csromer_b200.transformers.flaggers import hampel, median_absolute_deviation


@pytest.fixture(scope="module")
def g():
    return golden("next_rows")


def fresh(g):
    return cs.Dataset(nu=g["nu"], sigma=g["sigma"].copy(), data=g["data"].copy(), spectral_idx=0.7)


@pytest.mark.parametrize("tag,nsigma,delete", [("mean_ns0_w", 0.0, False), ("mean_ns5_w", 5.0, False), ("mean_ns5_del", 5.0, True)])
def test_mean_flagger_matches_reference(g, tag, nsigma, delete, capsys):
    ds = fresh(g)
    idxs, outliers = cs.MeanFlagger(dataset=ds, nsigma=nsigma, delete_channels=delete).run()
    np.testing.assert_array_equal(idxs, g[tag + "_idxs"])
    np.testing.assert_array_equal(outliers, g[tag + "_outliers"])
    np.testing.assert_array_equal(ds.w, g[tag + "_w"])
    np.testing.assert_array_equal(ds.lambda2, g[tag + "_lambda2"])
    np.testing.assert_allclose(ds.sigma, g[tag + "_sigma"], rtol=1e-14)
    np.testing.assert_array_equal(ds.data, g[tag + "_data"])
    assert ds.k == pytest.approx(float(g[tag + "_k"]), rel=1e-14)
    assert "Flagging" in capsys.readouterr().out


def test_hampel_and_manual_flaggers_match_reference(g):
    keep, outliers = hampel(g["sigma"], 9, 3.0, False)
    np.testing.assert_array_equal(keep, g["hampel_keep"])
    np.testing.assert_array_equal(outliers, g["hampel_outliers"])
    imputed, _, _ = hampel(g["sigma"], 9, 3.0, True)
    np.testing.assert_allclose(imputed, g["hampel_imputed"], rtol=1e-15)
    assert median_absolute_deviation(g["sigma"]) == pytest.approx(float(g["mad"]), rel=1e-15)
    # the class (the reference's run() cannot execute: it passes dataset.w as the window)
    ds = fresh(g)
    k_before = ds.k
    keep2, out2 = cs.HampelFlagger(dataset=ds, nsigma=3.0, w=9).run()
    np.testing.assert_array_equal(out2, g["hampel_outliers"])
    assert np.all(ds.w[out2] == 0) and ds.k == k_before          # in-place flagging leaves k stale, like the reference
    ds = fresh(g)
    idxs, _ = cs.ManualFlagger(dataset=ds, delete_channels=True, outlier_idxs=[3, 4, 50]).run()
    np.testing.assert_array_equal(idxs, g["manual_idxs"])
    np.testing.assert_array_equal(ds.lambda2, g["manual_lambda2"])
    np.testing.assert_allclose(ds.sigma, g["manual_sigma"], rtol=1e-14)
    np.testing.assert_array_equal(ds.data, g["manual_data"])


def test_per_pixel_mean_flagging_extension():
    rng = np.random.RandomState(0)
    m, P = 50, 7
    sigma = 1e-3 * (1 + 0.1 * rng.uniform(size=(m, P)))
    sigma[5, 2] *= 10
    sigma[9, 6] *= 10
    ds = cs.Dataset(nu=np.linspace(1e9, 2e9, m), sigma=sigma, data=np.zeros((m, P), np.complex64))
    keep, mask = cs.MeanFlagger(dataset=ds, nsigma=5.0).run()
    assert mask.shape == (m, P) and mask[5, 2] and mask[9, 6]
    from csromer_b200.transformers.flaggers import normal_flagging
    for p in range(P):          # every pixel gets exactly the single-line-of-sight decision
        _, outl = normal_flagging(sigma[:, p], None, 5.0)
        np.testing.assert_array_equal(np.where(mask[:, p])[0], outl)
    np.testing.assert_array_equal(ds.w == 0, mask)
    with pytest.raises(ValueError):
        cs.MeanFlagger(dataset=ds, nsigma=5.0, delete_channels=True).run()


def test_calculate_noise_and_filter_cubes_match_reference(g):
    cube = g["cube"]
    np.testing.assert_allclose(cs.calculate_noise(image=cube), g["noise_full"], rtol=1e-6)
    np.testing.assert_allclose(cs.calculate_noise(image=cube, x0=2, xn=20, y0=1, yn=15), g["noise_window"], rtol=1e-6)
    assert cs.calculate_noise(image=cube[1]) == pytest.approx(float(g["noise_image"]), rel=1e-6)
    with pytest.raises(NotImplementedError):
        cs.calculate_noise(image=cube, use_sigma_clipped_stats=True)
    header = {"CRVAL3": 1.0e9, "NAXIS3": 6, "CDELT3": 1.0e6}
    _, fQ, _, fnu = cs.filter_cubes(np.zeros_like(cube), cube, cube * 0.5, header)
    np.testing.assert_array_equal(fQ, g["filter_Q"])
    np.testing.assert_array_equal(fnu, g["filter_nu"])


def test_fits_codec_layout_and_round_trip(tmp_path):
    rng = np.random.RandomState(1)
    cube = (rng.normal(size=(5, 4, 6)) + 1j * rng.normal(size=(5, 4, 6))).astype(np.complex64)
    phi = np.linspace(-100.0, 100.0, 5)
    header = fits_lite.Header({"CRVAL1": 12.5, "CTYPE1": "RA---SIN", "BUNIT": "Jy/beam", "EQUINOX": 2000.0})
    path = str(tmp_path / "fd_cube.fits")
    cs.Writer(output=path).writeFITSCube(cube, header, 5, phi, phi[1] - phi[0])
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T"
    hdr, data = cs.Reader().readCube(file=path, memmap=True)
    assert data.shape == (2, 5, 4, 6) and hdr["NAXIS"] == 4 and hdr["NAXIS4"] == 2 and hdr["NAXIS3"] == 5
    assert hdr["CTYPE3"] == "Phi" and hdr["CUNIT3"] == "rad/m/m" and hdr["CRVAL3"] == -100.0 and hdr["CDELT3"] == 50.0
    assert hdr["CTYPE1"] == "RA---SIN" and hdr["BUNIT"] == "Jy/beam" and hdr["CRVAL1"] == 12.5
    np.testing.assert_array_equal(np.asarray(data[0]), cube.real)
    np.testing.assert_array_equal(np.asarray(data[1]), cube.imag)
    # big-endian float32 right after the header block(s)
    first = np.frombuffer(raw[2880:2884] if raw[2880:2881] != b" " else raw[5760:5764], dtype=">f4")[0]
    assert first == cube.real[0, 0, 0]
    # plain image, integer type, non-memmap
    img = np.arange(12, dtype=np.int16).reshape(3, 4)
    p2 = str(tmp_path / "img.fits")
    cs.Writer().writeFITS(data=img, header={"OBJECT": "test"}, output=p2)
    h2, d2 = cs.Reader().readImage(name=p2)
    assert h2["BITPIX"] == 16 and h2["OBJECT"] == "test"
    np.testing.assert_array_equal(d2, img)
    assert cs.Reader(Q_cube_name=p2).readHeader()["NAXIS1"] == 4
    # numpy side of the reader / writer
    npy = str(tmp_path / "qu.npy")
    cs.Writer(output=npy).writeNPCube(np.stack([img, 2 * img], axis=-1))
    Q, U = cs.Reader(numpy_file=npy).readNumpyFile()
    np.testing.assert_array_equal(U, 2 * Q)


def test_reconstructor_wrapper_static_helpers():
    W = cs.CSROMERReconstructorWrapper
    # 3-point parabola: exact for a sampled parabola
    n, cell = 64, 2.5
    j = np.arange(n)
    true_pos, true_peak = 40.3, 7.0
    f = (true_peak - 0.5 * (j - true_pos) ** 2).clip(min=0.0) * np.exp(1j * 0.3)
    pos, peak = W.estimate_peak_quadratic_interpolation(f, cell)
    assert pos == pytest.approx((true_pos - n / 2) * cell, abs=1e-9) and peak == pytest.approx(true_peak, abs=1e-9)
    pos2, peak2 = W.estimate_peak_quadratic_interpolation(np.stack([f, np.roll(f, 3)], axis=1), cell)   # batched
    assert pos2[0] == pytest.approx(pos) and pos2[1] == pytest.approx(pos + 3 * cell) and peak2[1] == pytest.approx(peak)
    assert W.calculate_ricean_peak(5.0, 1.0) == pytest.approx(np.sqrt(25 - 2.3))
    assert W.calculate_sigma_phi_peak(50.0, 10.0, 0.5) == pytest.approx(50.0 / (2 * 10.0 / 0.5))
    # sigma clipping as documented by astropy: outliers beyond sigma * mad_std around the mean are dropped iteratively
    from csromer_b200.wrappers.reconstructors import sigma_clipped_std
    rng = np.random.RandomState(0)
    x = rng.normal(0, 1.0, 4000)
    x[:20] = 50.0
    assert sigma_clipped_std(x, sigma=3.0) == pytest.approx(np.std(x[20:][np.abs(x[20:]) < 3.05]), rel=0.02)
    assert sigma_clipped_std(x, mask=np.arange(4000) >= 20, sigma=3.0) == 0.0          # only the 20 equal outliers are left
    # the reference's default threshold = 0 masks everything but phi = 0 -> noise 0 (csromer_reconstructor.py:86-102)
    phi = np.linspace(-10, 10, 21)
    assert W.calculate_fd_signal_noise(rng.normal(size=21) + 1j * rng.normal(size=21), phi, 10.0) == 0.0
    assert W.calculate_fd_signal_noise(rng.normal(size=21) + 1j * rng.normal(size=21), phi, 10.0, threshold=2.0, sigma=3.0) > 0


@pytest.mark.gpu
def test_reconstructor_wrapper_matches_oracle_pipeline(capsys):
    """the reference's one-call wrapper (default lambda / noise schedule, flagger first) on the GPU path against the
    oracle run with the same schedule; step = 1 / lambda_max on both sides (the unit step is unstable at the end of
    this schedule)"""
    from oracle import restatement as ob
    rng = np.random.RandomState(4)
    m = 160
    nu = np.linspace(1.008e9, 2.031e9, m)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-150.0, spectral_idx=1.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    sig = np.asarray(src.sigma).copy()
    sig[[11, 70]] *= 8.0
    src.sigma = sig
    flagger = cs.MeanFlagger(dataset=src, nsigma=5.0, delete_channels=True)
    rec = cs.CSROMERReconstructorWrapper(dataset=src, flagger=flagger, oversampling=8.0, step="auto")
    rec.reconstruct()
    capsys.readouterr()
    assert src.m == m - 2 and rec.fd_model.shape == (len(rec.parameter.phi),)
    par = rec.parameter
    op = ob.ExactDFT(src.lambda2, par.phi, src.w, src.s, src.k, src.l2_ref)
    dirty = op.backward(np.asarray(src.data))
    lam0 = np.sqrt(src.m + 2 * np.sqrt(src.m)) * np.sqrt(2) * np.mean(src.sigma)
    assert rec.lambda_l_norm == pytest.approx(lam0, rel=1e-12)
    step = 1.0 / cs.estimate_lipschitz(rec.dft)
    ref = ob.fista(op, np.asarray(src.data), ob.complex_to_real(dirty), lam0, src.theo_noise, None, step=step)
    model = ob.real_to_complex(ref["x"])
    scale = np.abs(model).max()
    assert np.abs(rec.fd_dirty - dirty).max() / np.abs(dirty).max() < 1e-5
    assert np.abs(rec.fd_model - model).max() / scale < 2e-4
    assert rec.rm_model == par.phi[np.argmax(np.abs(model))] and abs(rec.rm_model + 150.0) < 2 * par.cellsize
    mag = np.abs(model)
    first = np.sum(par.phi * mag) / mag.sum()
    assert rec.second_moment == pytest.approx(np.sum(mag * (par.phi - first) ** 2) / mag.sum(), rel=2e-3)
    resid_fd = op.backward(np.asarray(src.data) - ref["model"])
    assert np.abs(rec.fd_residual - resid_fd).max() / np.abs(dirty).max() < 2e-4
    restored = ob.convolve_beam(model, ob.gaussian_beam(par.rmtf_fwhm, par.cellsize)) + resid_fd
    assert np.abs(rec.fd_restored - restored).max() / np.abs(restored).max() < 2e-4
    assert abs(rec.rm_restored_quadratic_interpolation + 150.0) < par.cellsize


@pytest.mark.gpu
def test_device_cube_transposes_and_channel_statistics():
    import torch
    rng = np.random.RandomState(5)
    m, Y, X = 70, 37, 45          # nothing is a multiple of the 32 x 32 tile
    Q = rng.normal(size=(m, Y, X)).astype(np.float32)
    U = rng.normal(size=(m, Y, X)).astype(np.float32)
    Q[3, 5, 7] = np.nan
    U[10, 0, 0] = np.inf
    Q[20] = np.nan
    cube = cs.DeviceCube(Q, U)
    npix = Y * X
    for p0, P in ((0, 33), (100, 257), (npix - 50, 50)):
        planar, valid = cube.slab(p0, P)
        qf, uf = Q.reshape(m, -1)[:, p0:p0 + P].T, U.reshape(m, -1)[:, p0:p0 + P].T
        ok = np.isfinite(qf) & np.isfinite(uf)
        np.testing.assert_array_equal(valid.cpu().numpy().astype(bool), ok)
        np.testing.assert_array_equal(planar[:, :m].cpu().numpy(), np.where(ok, qf, 0))
        np.testing.assert_array_equal(planar[:, m:2 * m].cpu().numpy(), np.where(ok, uf, 0))
        assert planar.shape[1] % 8 == 0 and float(planar[:, 2 * m:].abs().sum()) == 0
    s_sum, s_std = cube.channel_stats("Q", x0=2, xn=40, y0=1, yn=30)
    want_sum, want_std = np.nansum(Q[:, 1:30, 2:40], axis=(1, 2)), np.nanstd(Q[:, 1:30, 2:40].astype(np.float64), axis=(1, 2))
    fin = np.isfinite(want_std)
    np.testing.assert_allclose(s_sum.cpu().numpy()[fin], want_sum[fin], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(s_std.cpu().numpy()[fin], want_std[fin], rtol=1e-6)
    assert np.isnan(s_std.cpu().numpy()[20]) and not fin[20]
    np.testing.assert_allclose(cs.calculate_noise(image=torch.as_tensor(U[:, :, :]).cuda()).cpu().numpy()[1:10],
                               np.nanstd(U[1:10].astype(np.float64), axis=(1, 2)), rtol=1e-6)
    # Faraday-depth side: planar rows -> (2, n, Y, X)
    n = 53
    out = cube.new_output(n)
    F = rng.normal(size=(npix, 2 * n + 2)).astype(np.float32)
    for p0 in range(0, npix, 400):
        P = min(400, npix - p0)
        cube.store(torch.as_tensor(F[p0:p0 + P]).cuda(), n, p0, out)
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got[0].reshape(n, -1), F[:, :n].T)
    np.testing.assert_array_equal(got[1].reshape(n, -1), F[:, n:2 * n].T)


# ---- device flaggers and the Galactic RM foreground ---------------------------------
@pytest.mark.gpu
def test_device_flaggers_match_the_reference_per_pixel():
    """csr_flag_mean / csr_flag_hampel on a [P, m] sigma: every row that carries the golden sigma array reproduces the
    reference's own outlier lists (tests/golden/next_rows.npz), random rows reproduce the host mirrors of the reference
    functions, and MeanFlagger / HampelFlagger use the kernels for per-pixel sigma."""
    import torch

    import csromer_b200 as cs
    from csromer_b200.transformers.flaggers.flagger import device_flag
    from csromer_b200.transformers.flaggers.hampel_flagger import hampel
    from csromer_b200.transformers.flaggers.mean_flagger import normal_flagging
    g = golden("next_rows")
    sigma = g["sigma"]
    m, P = sigma.shape[0], 37
    rng = np.random.RandomState(9)
    sig = np.repeat(sigma[:, None], P, axis=1)
    sig[:, 1::2] = 0.001 * (1.0 + 0.2 * rng.uniform(size=(m, P // 2)))      # odd pixels: their own noise levels
    sig[rng.randint(0, m, 40), rng.randint(0, P // 2, 40) * 2 + 1] *= 5.0
    for nsig, key in ((0.0, "mean_ns0_w_outliers"), (5.0, "mean_ns5_w_outliers")):
        w, mask = device_flag(sig, np.ones(m), "mean", nsig)
        for p in range(P):
            expect = g[key] if p % 2 == 0 else normal_flagging(sig[:, p], None, nsig)[1]
            np.testing.assert_array_equal(np.nonzero(mask[:, p])[0], expect)
            assert (w[mask[:, p], p] == 0).all() and (w[~mask[:, p], p] == 1).all()
    w, mask = device_flag(sig, np.ones(m), "hampel", 3.0, window=9)
    for p in range(P):
        expect = g["hampel_outliers"] if p % 2 == 0 else hampel(sig[:, p], 9, 3.0, False)[1]
        np.testing.assert_array_equal(np.nonzero(mask[:, p])[0], expect)
    # through the classes (zero-weighting of a per-pixel dataset), CUDA tensors stay on the device
    ds = cs.Dataset(nu=g["nu"], sigma=sig, spectral_idx=0.7)
    keep, out = cs.MeanFlagger(dataset=ds, nsigma=5.0).run()
    np.testing.assert_array_equal(np.nonzero(out[:, 0])[0], g["mean_ns5_w_outliers"])
    assert (ds.w[out] == 0).all() and (ds.w[~out] > 0).all()
    ds = cs.Dataset(nu=g["nu"], sigma=sig, spectral_idx=0.7)
    keep, out = cs.HampelFlagger(dataset=ds, nsigma=3.0, w=9).run()
    np.testing.assert_array_equal(np.nonzero(out[:, 0])[0], g["hampel_outliers"])
    wt, mt = device_flag(torch.as_tensor(sig, dtype=torch.float32).cuda(), torch.ones(m).cuda(), "mean", 5.0)
    assert wt.is_cuda and mt.is_cuda and bool((mt.cpu().numpy() == device_flag(sig, np.ones(m), "mean", 5.0)[1]).all())


@pytest.mark.gpu
def test_galactic_rm_image_and_derotation():
    """FaradaySky on a synthetic HEALPix map (value = pixel index: the gathered map IS the index chain): image lookup
    against the oracle's restatement, both orderings, the reference's transposed argument order, position lookup, and the
    result feeding Dataset.subtract_galacticrm's device twin."""
    import csromer_b200 as cs
    from csromer_b200 import _device as dv
    nside = 64
    npix = 12 * nside * nside
    sky = cs.FaradaySky(data=(np.arange(npix, dtype=np.float32), 0.5 * np.arange(npix, dtype=np.float32)))
    assert sky.nside == nside and sky.ordering == "ring"
    for ctype, crval in (("RA---SIN", (150.0, -30.0)), ("RA---TAN", (10.0, 62.0)), ("RA---TAN", (266.4, -28.9))):
        header = {"NAXIS1": 40, "NAXIS2": 28, "CRVAL1": crval[0], "CRVAL2": crval[1], "CRPIX1": 20.5, "CRPIX2": 14.0,
                  "CDELT1": -0.25, "CDELT2": 0.25, "CTYPE1": ctype, "CTYPE2": ctype.replace("RA--", "DEC-"), "RADESYS": "ICRS"}
        mean, std = sky.galactic_rm_image(header)
        ref_mean, ref_std = ob.galactic_rm_image(np.arange(npix), 0.5 * np.arange(npix), nside, header)
        assert mean.shape == (28, 40)
        np.testing.assert_array_equal(mean.astype(np.int64), ref_mean)
        np.testing.assert_array_equal(std, 0.5 * ref_mean)
        assert len(np.unique(mean)) > 30            # the 10 x 7 degree field really crosses many map pixels
    square = dict(header, NAXIS1=28)
    swapped, _ = sky.galactic_rm_image(square, reference_axis_order=True)
    np.testing.assert_array_equal(swapped, sky.galactic_rm_image(square)[0].T)
    # positions: pixel centres map to themselves, in both frames and both orderings
    pix = np.random.RandomState(2).randint(0, npix, 500)
    theta, phi = ob.healpix_pix2ang_ring(nside, pix)
    got, _ = sky.galactic_rm(ra=phi, dec=0.5 * np.pi - theta, frame="galactic")
    np.testing.assert_array_equal(got.astype(np.int64), pix)
    ra = np.radians(np.array([192.85948, 266.40499])), np.radians(np.array([27.12825, -28.93617]))
    l, b = ob.icrs_to_galactic(*ra)
    got, _ = sky.galactic_rm(ra=ra[0], dec=ra[1], frame="icrs")
    np.testing.assert_array_equal(got.astype(np.int64), ob.healpix_ang2pix_ring(nside, 0.5 * np.pi - b, l))
    # NESTED: the same sky position must land in the pixel whose nested index corresponds to the ring one
    lib = cs._lib.load()
    import torch
    lon, lat = torch.as_tensor(phi).cuda(), torch.as_tensor(0.5 * np.pi - theta).cuda()
    pr, pn = torch.empty(500, dtype=torch.int64).cuda(), torch.empty(500, dtype=torch.int64).cuda()
    for nested, out in ((0, pr), (1, pn)):
        cs._lib.check(lib.csr_healpix_lookup(None, None, nside, nested, dv.tptr(lon), dv.tptr(lat), 0, 500, None, None, dv.tptr(out),
                                             dv.stream_ptr()), "csr_healpix_lookup")
    np.testing.assert_array_equal(pr.cpu().numpy(), pix)
    pn = pn.cpu().numpy()
    assert len(np.unique(pn)) == len(np.unique(pix)) and (pn // (nside * nside) < 12).all()
    north = np.cos(theta) > 2 / 3
    assert (pn[north] // (nside * nside) < 4).all() and (pn[np.cos(theta) < -2 / 3] // (nside * nside) >= 8).all()
    # the per-pixel foreground feeds the de-rotation (base/dataset.py:377-381)
    nu = np.linspace(1.0e9, 2.0e9, 64)
    header = {"NAXIS1": 6, "NAXIS2": 5, "CRVAL1": 150.0, "CRVAL2": -30.0, "CRPIX1": 3.0, "CRPIX2": 3.0, "CDELT1": -2.0,
              "CDELT2": 2.0, "CTYPE1": "RA---SIN"}
    small = cs.FaradaySky(data=(np.random.RandomState(3).normal(0, 30, npix).astype(np.float32), np.ones(npix, np.float32)))
    rm, _ = small.galactic_rm_image(header)
    src = cs.FaradayThinSource(nu=nu, s_nu=1.0, phi_gal=rm.reshape(-1).astype(np.float64), spectral_idx=0.0)
    src.simulate()
    before = np.asarray(src.data).copy()
    src.subtract_galacticrm(rm.reshape(-1))
    assert np.abs(np.asarray(src.data) - 1.0).max() < 1e-6 and np.abs(before - 1.0).max() > 0.5
