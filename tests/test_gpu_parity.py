"""GPU parity: the CUDA path (through the Python mirror of the reference API -> ctypes -> C ABI ->
libcsromer_b200.so) against (a) the golden fixtures produced by the reference's own classes and (b) the float64
oracle on seeded inputs.

Stated tolerances (max-norm, relative to the peak of the reference quantity; north_star / SURVEY.md section 8c):
    TOL_TRANSFORM = 1e-5   single transforms  F(phi), P(lambda^2), RMTF
    TOL_MODEL     = 1e-4   reconstructed model / solution after K FISTA iterations
Observed on B200: ~2e-6 and ~1e-6.  Channels with w == 0 are excluded from model_data comparisons (the reference
leaves them uninitialised).
"""
import numpy as np
import pytest
import torch
from conftest import golden

import csromer_b200 as cs
from oracle import restatement as ob

pytestmark = pytest.mark.gpu

TOL_TRANSFORM = 1e-5
TOL_MODEL = 1e-4


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module", autouse=True)
def _library(built_library):
    assert torch.cuda.is_available(), "run the gpu-marked tests on a B200"
    yield


def source_from_golden(g):
    """a Dataset carrying exactly the arrays the reference used"""
    ds = cs.Dataset(lambda2=g["lambda2"], data=g["data"], sigma=g["sigma"], spectral_idx=float(g["spectral_idx"]))
    par = cs.Parameter()
    par.calculate_cellsize(dataset=ds, oversampling=8, verbose=False)
    np.testing.assert_array_equal(par.phi, g["phi"])
    return ds, par


# ---------------------------------------------------------------------------------------------------------
# (a) against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["thin_nonoise_jvla1000", "mixed_noise1_jvla1000_k30", "mixed_noise1_jvla200_k60",
                                  "thin_noise2_meerkat300_k40"])
def test_transforms_match_reference(name):
    g = golden(name)
    ds, par = source_from_golden(g)
    dft = cs.NDFT1D(dataset=ds, parameter=par)
    F = dft.backward(ds.data)
    assert F.dtype == np.complex64 and F.shape == (int(g["n"]),)
    assert rel(F, g["F_dirty"]) < TOL_TRANSFORM
    assert rel(dft.RMTF(), g["rmtf"]) < TOL_TRANSFORM
    assert rel(dft.forward(g["x_test"]), g["fwd_x_test"]) < TOL_TRANSFORM
    assert rel(dft.forward_normalized(g["x_test"]), g["fwdn_x_test"]) < TOL_TRANSFORM
    # the README's alias and the NUFFT1D API evaluate the same operator
    assert cs.DFT1D is cs.NDFT1D
    nufft = cs.NUFFT1D(dataset=ds, parameter=par, solve=True)
    assert rel(nufft.backward(ds.data), g["F_dirty"]) < TOL_TRANSFORM
    assert rel(nufft.forward_normalized(g["x_test"]), g["fwdn_x_test"]) < TOL_TRANSFORM


def test_thin_source_known_answers():
    g = golden("thin_nonoise_jvla1000")
    ds, par = source_from_golden(g)
    dft = cs.NDFT1D(dataset=ds, parameter=par)
    F = dft.backward(ds.data)
    assert par.phi[np.argmax(np.abs(F))] == pytest.approx(-201.9582742926516)
    assert np.abs(F).max() == pytest.approx(0.003354779466042806, rel=1e-5)
    R = dft.RMTF()
    assert np.abs(R).max() == pytest.approx(1.0, abs=1e-5) and par.phi[np.argmax(np.abs(R))] == 0.0


@pytest.mark.parametrize("name", ["thin_noise0_jvla1000_k100", "mixed_noise1_jvla1000_k30", "mixed_noise1_jvla200_k60",
                                  "thin_noise2_meerkat300_k40"])
@pytest.mark.parametrize("operator", ["NDFT1D", "NUFFT1D"])
def test_fista_matches_reference(name, operator):
    """the README recipe (README.md:157-210), delta basis, against the reference's own run"""
    g = golden(name)
    ds, par = source_from_golden(g)
    dft = cs.NDFT1D(dataset=ds, parameter=par)
    op = dft if operator == "NDFT1D" else cs.NUFFT1D(dataset=ds, parameter=par, solve=True)
    par.data = dft.backward(ds.data)
    par.complex_data_to_real()
    chi2 = cs.Chi2(dft_obj=op, wavelet=None)
    l1 = cs.L1(reg=float(g["lam0"]))
    F_obj, g_obj = cs.OFunction([chi2, l1]), cs.OFunction([l1])
    opt = cs.FISTA(guess_param=par, F_obj=F_obj, fx=chi2, gx=g_obj, noise=float(g["noise"]), maxiter=int(g["maxiter"]),
                   verbose=False)
    obj, X = opt.run()
    assert X is not par and X.data.shape == (2 * int(g["n"]),)
    assert np.isfinite(X.data).all()
    assert rel(X.data, g["fista_x"]) < TOL_MODEL
    assert rel(ds.model_data, g["model_data"]) < TOL_MODEL
    assert rel(ds.residual, g["residual"]) < TOL_MODEL
    assert l1.reg == pytest.approx(float(g["lam_final"]), rel=1e-12)  # lambda decayed in place (fista.py:100-102)
    assert obj == pytest.approx(float(g["fista_obj"]), rel=1e-4)
    assert np.count_nonzero(X.data) == np.count_nonzero(g["fista_x"])
    X.real_data_to_complex()
    assert X.data.shape == (int(g["n"]),)
    assert rel(chi2.F_dirty, g["F_dirty"]) < TOL_TRANSFORM


# ---------------------------------------------------------------------------------------------------------
# (b) seeded batches against the oracle
# ---------------------------------------------------------------------------------------------------------
def make_batch(nch, P, seed, per_pixel_w=False, flags=0.0, l2_ref=None, mixed=False):
    rng = np.random.RandomState(seed)
    nu = np.linspace(1.008e9, 2.031e9, nch)
    kwargs = dict(nu=nu, spectral_idx=1.0)
    if l2_ref is not None:
        kwargs["l2_ref"] = l2_ref
    src = cs.FaradayThinSource(s_nu=0.0035 * rng.uniform(0.5, 1.5, size=P), phi_gal=rng.uniform(-500, 500, size=P),
                               **kwargs)
    src.simulate()
    if mixed:
        thick = cs.FaradayThickSource(s_nu=0.0035, phi_fg=rng.uniform(50, 200, size=P), phi_center=200.0, **kwargs)
        thick.simulate()
        src.data = src.data + thick.data * (rng.uniform(size=P) < 0.3)
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    if per_pixel_w:
        src.sigma = 0.0007 * rng.uniform(0.5, 2.0, size=(nch, P))
        if flags > 0:
            w = src.w.copy()
            w[rng.uniform(size=w.shape) < flags] = 0.0
            src.w = w
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    return src, par


def oracle_operator(src, par):
    return ob.ExactDFT(src.lambda2, par.phi, src.w, src.s, src.k, src.l2_ref)


@pytest.mark.parametrize("nch,P,kw", [
    (200, 1, {}), (200, 127, {}), (200, 129, {}), (300, 261, {}), (203, 33, {}),  # ragged tiles, m not a multiple of 4
    (200, 70, dict(per_pixel_w=True, flags=0.2)), (200, 50, dict(l2_ref=0.04)),
])
def test_batched_transforms(nch, P, kw):
    src, par = make_batch(nch, P, seed=100 + P, **kw)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    op = oracle_operator(src, par)
    data = np.asarray(src.data)
    F = dft.backward(data)
    assert F.shape == (len(par.phi), P)
    assert rel(F, op.backward(data)) < TOL_TRANSFORM
    rng = np.random.RandomState(1)
    x = (rng.normal(size=(len(par.phi), P)) + 1j * rng.normal(size=(len(par.phi), P))) * (rng.uniform(size=(len(par.phi), P)) < 0.02)
    good = np.asarray(src.w) > 0
    good = good if good.ndim == 2 else good[:, None] & np.ones((1, P), bool)
    assert rel(dft.forward(x), op.forward(x)) < TOL_TRANSFORM
    fn, fo = dft.forward_normalized(x), op.forward_normalized(x)
    assert rel(fn[good], fo[good]) < TOL_TRANSFORM and np.all(fn[~good] == 0)
    assert rel(dft.RMTF(), op.rmtf()) < TOL_TRANSFORM


def run_fista(src, par, K, wav=None, lam_mult=40.0, step=1.0, fused=None, x0="dirty", maxiter="K"):
    dft = cs.NDFT1D(dataset=src, parameter=par)
    lam0 = lam_mult * np.asarray(src.theo_noise)
    noise = lam0 / (2.0 * K)
    guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
    guess.data = dft.backward(src.data) if x0 == "dirty" else np.zeros_like(dft.backward(src.data))
    guess.complex_data_to_real()
    if wav is not None:
        guess.data = wav.decompose(guess.data)
    chi2 = cs.Chi2(dft_obj=dft, wavelet=wav)
    l1 = cs.L1(reg=lam0)
    opt = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                   maxiter=K if maxiter == "K" else maxiter, verbose=False, step=step, fused=fused)
    cost, X = opt.run()
    return cost, X, l1, lam0, noise


def oracle_fista(src, par, K, lam0, noise, owav=None, step=1.0, x0="dirty", maxiter="K"):
    op = oracle_operator(src, par)
    data = np.asarray(src.data)
    start = ob.complex_to_real(op.backward(data))
    dirty_peak = np.abs(start).max()
    if x0 != "dirty":
        start = np.zeros_like(start)
    if owav is not None:
        start = owav.decompose(start)
    ref = ob.fista(op, data, start, lam0, noise, K if maxiter == "K" else maxiter, wavelet=owav, step=step)
    # parity on a diverging run is meaningless: insist the oracle itself stayed bounded
    assert np.isfinite(ref["x"]).all() and np.abs(ref["x"]).max() < (10 if owav is None else 100) * dirty_peak
    return ref


def check_fista(src, X, cost, l1, ref):
    good = np.asarray(src.w) > 0
    if good.ndim == 1 and np.ndim(src.data) == 2:
        good = good[:, None] & np.ones((1, np.shape(src.data)[1]), bool)
    ex, em, ec = rel(X.data, ref["x"]), rel(np.asarray(src.model_data)[good], ref["model"][good]), rel(cost, ref["cost"])
    assert ex < TOL_MODEL and em < TOL_MODEL and ec < TOL_MODEL, (ex, em, ec)
    np.testing.assert_allclose(l1.reg, ref["lam_final"], rtol=1e-9)
    # identical support except where |u| - lambda is within rounding of zero
    mism = np.count_nonzero((np.asarray(X.data) != 0) != (ref["x"] != 0))
    assert mism <= max(2, int(1e-4 * ref["x"].size)), mism


@pytest.mark.parametrize("nch,P,K,kw", [
    (200, 5, 30, {}), (200, 261, 40, dict(mixed=True)), (200, 140, 40, dict(per_pixel_w=True, flags=0.15)),
    (300, 64, 25, dict(l2_ref=0.05)), (1000, 40, 50, dict(mixed=True)),
])
def test_fista_delta_basis_batches(nch, P, K, kw):
    src, par = make_batch(nch, P, seed=7 + P, **kw)
    cost, X, l1, lam0, noise = run_fista(src, par, K)
    ref = oracle_fista(src, par, K, lam0, noise)
    assert X.data.shape == (2 * len(par.phi), P) and np.shape(cost) == (P,)
    check_fista(src, X, cost, l1, ref)


# (m = 200, 1000: the direct row-contiguous forward epilogue; 202: m % 4 != 0, sparse tiles stay on the pipelined fragment epilogue;
#  203: odd m, the general epilogue; 204 with per-pixel weights: channel factors of the pixel's own in the direct epilogue)
@pytest.mark.parametrize("nch,P,K,kw", [(200, 261, 30, dict(mixed=True)), (1000, 300, 20, {}), (203, 129, 25, dict(per_pixel_w=True, flags=0.1)),
                                        (202, 140, 20, {}), (204, 150, 20, dict(per_pixel_w=True, flags=0.15))])
def test_sparsity_shortcuts_are_bit_exact(nch, P, K, kw):
    """The epilogue's all-zero-block shortcut, the forward GEMM's K-block skipping and the one-product screening of
    the adjoint must not change a single bit."""
    from csromer_b200 import _lib
    src, par = make_batch(nch, P, seed=11 + P, **kw)
    results = []
    try:
        # (sym: the pair kernels screen +phi and -phi together, half the K; pair = 1, sym = 0 is the plain pair kernel)
        # (incr: incremental screening -- a tile that passed is tested again only when its margin may be used up)
        # (pair3 / tma: the three-product transforms on CTA pairs, with the residual epilogue staged through shared memory)
        # (grp: sparse tiles of the forward transform accumulate their active K blocks as one chunk -- the one switch here that
        #  regroups partial sums, so the runs form two families, with and without it; the forward kernels on 128 x 128 tiles
        #  (f128) and on CTA pairs (pair3 = 2) do not group and belong to the second)
        for zero_skip, k_skip, screen, screen8, pair, sym, incr, pair3, tma, f128, grp in (
                (1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 16), (1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 16), (1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 16),
                (1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 16), (1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 16), (1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 16),
                (1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 16), (0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 16), (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 16),
                (1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0), (1, 1, 1, 1, 1, 1, 1, 2, 1, 0, 16), (1, 1, 1, 1, 1, 1, 1, 2, 0, 0, 16),
                (1, 1, 1, 1, 1, 0, 0, 2, 1, 0, 16), (1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 16), (0, 0, 0, 0, 1, 0, 0, 2, 1, 0, 16),
                (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)):
            _lib.set_option("screen_incr", incr)
            _lib.set_option("rescale", incr)
            _lib.set_option("pair3", pair3)
            _lib.set_option("tma_epi", tma)
            _lib.set_option("fwd128", f128)
            _lib.set_option("fwd_group", grp)
            _lib.set_option("zero_skip", zero_skip)
            _lib.set_option("k_skip", k_skip)
            _lib.set_option("screen", screen)
            _lib.set_option("screen8", screen8)
            _lib.set_option("pair", pair)
            _lib.set_option("sym", sym)
            _lib.set_option("resid_fast", zero_skip and screen8)   # (the lean residual epilogue: off in the plainer combinations)
            cost, X, l1, lam0, noise = run_fista(src, par, K)
            grouped = grp > 0 and f128 == 0 and pair3 < 2
            results.append((grouped, np.array(X.data, copy=True), np.array(cost, copy=True), np.array(src.model_data, copy=True)))
    finally:
        _lib.set_option("zero_skip", 1)
        _lib.set_option("k_skip", 1)
        _lib.set_option("screen", 1)
        _lib.set_option("screen8", 1)
        _lib.set_option("pair", 1)
        _lib.set_option("sym", 1)
        _lib.set_option("resid_fast", 1)
        _lib.set_option("screen_incr", 1)
        _lib.set_option("rescale", 1)
        _lib.set_option("pair3", 1)
        _lib.set_option("tma_epi", 1)
        _lib.set_option("fwd128", 0)
        _lib.set_option("fwd_group", 16)
    assert np.count_nonzero(results[0][1]) > 0 and np.count_nonzero(results[0][1]) < 0.2 * results[0][1].size
    for family in (True, False):
        runs = [r[1:] for r in results if r[0] == family]
        assert len(runs) >= 7
        for other in runs[1:]:
            for a, b in zip(runs[0], other):
                np.testing.assert_array_equal(a, b)
    # the two families differ in rounding only
    assert rel(results[0][1], results[-1][1]) < 1e-5
    ref = oracle_fista(src, par, K, lam0, noise)
    check_fista(src, X, cost, l1, ref)


@pytest.mark.parametrize("nch,P,K,kw", [(200, 300, 30, dict(mixed=True)), (1000, 261, 20, dict(mixed=True)),
                                        (204, 290, 25, dict(per_pixel_w=True, flags=0.15))])
def test_forward_tiles_on_sixteen_epilogue_warps_keep_the_bits(nch, P, K, kw):
    """option fwd16: the sparse forward tiles run on dft_fwd_direct_kernel (sixteen epilogue warps, 64 columns per warp), the
    rest on dft_gemm_kernel<EPI_RESID>: same element arithmetic, same solution bit for bit"""
    from csromer_b200 import _lib
    src, par = make_batch(nch, P, seed=90 + P, **kw)
    out = []
    try:
        for fwd16 in (0, 1):
            _lib.set_option("fwd16", fwd16)
            cost, X, l1, lam0, noise = run_fista(src, par, K)
            out.append((np.array(X.data, copy=True), np.array(cost, copy=True), np.array(src.model_data, copy=True)))
    finally:
        _lib.set_option("fwd16", 1)
    assert 0 < np.count_nonzero(out[0][0]) < 0.2 * out[0][0].size
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


def test_fista_reuses_the_dirty_spectrum_it_was_started_from():
    """README.md:183-190 with torch data: F_dirty = dft.backward(dataset.data) becomes the guess.  When the guess still is that
    untouched tensor, the fused loop takes the planar device copy `backward` kept (no transpose back, no second dirty spectrum);
    a guess of the same VALUES without that provenance takes the general path.  Same solution either way, and a modified
    guess is honoured."""
    import torch
    src, par = make_batch(200, 140, seed=61, mixed=True)
    K = 20
    lam0 = 40.0 * np.asarray(src.theo_noise)
    noise = lam0 / (2.0 * K)
    results = []
    for variant in ("provenance", "clone", "modified"):
        ds = cs.Dataset(lambda2=src.lambda2, data=torch.as_tensor(np.asarray(src.data)), sigma=src.sigma, spectral_idx=1.0)
        dft = cs.NDFT1D(dataset=ds, parameter=par)
        f_dirty = dft.backward(ds.data)
        guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
        guess.data = f_dirty
        guess.complex_data_to_real()
        assert dft.dirty_of(guess.complex_source()) is not None
        if variant == "clone":
            guess.data = guess.data.clone()
            assert guess.complex_source() is None
        elif variant == "modified":
            guess.data[:, 0] = 0.0            # in place: the version counter of the tensor moves
            assert guess.complex_source() is None
        chi2, l1 = cs.Chi2(dft_obj=dft, F_dirty=f_dirty), cs.L1(reg=lam0.copy())
        opt = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise, maxiter=K,
                       verbose=False)
        cost, X = opt.run()
        reused = opt.device_outputs["f_dirty"] is dft._dirty_cache["planar"]
        assert reused == (variant == "provenance")
        results.append((X.data.cpu().numpy(), np.asarray(cost.cpu() if hasattr(cost, "cpu") else cost), ds.model_data.cpu().numpy()))
    for a, b in zip(results[0], results[1]):
        assert rel(a, b) < 2e-6
    assert rel(results[0][0][:, 1:], results[2][0][:, 1:]) < 2e-6 and rel(results[0][0][:, 0], results[2][0][:, 0]) > 1e-3
    ref = oracle_fista(src, par, K, lam0, noise)
    assert rel(results[0][0], ref["x"]) < TOL_MODEL


def test_repeated_call_is_replayed_as_a_cuda_graph():
    """csr_fista called again with identical arguments (same buffers: `out=` reuse) is captured on the second call and
    replayed as one CUDA graph from then on: same bits as plain launches, same launch accounting, new data honoured."""
    import torch
    from csromer_b200 import _device as dv
    from csromer_b200 import _lib
    src, par = make_batch(200, 150, seed=77, mixed=True)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    plan = dft.device_plan()
    dev = plan.device
    data, _ = dv.pack_planar(src.data, plan.m, plan.ld2m, dev)
    P, K = data.shape[0], 25
    w = torch.as_tensor(np.asarray(src.w), dtype=torch.float32, device=dev)
    s = torch.as_tensor(np.asarray(src.s), dtype=torch.float32, device=dev)
    k = dv.per_pixel_vector(src.k, P, dev, "k")
    lam = torch.full((P,), 40.0 * float(np.mean(src.theo_noise)), device=dev)
    nz = lam / (2 * K)
    lib = _lib.load()

    def run(x, out=None):
        lib.csr_launch_count(1)
        o = plan.fista(data, w, s, k, x, lam, nz, K, use_dirty_as_guess=True, want_delta=True, out=out)
        torch.cuda.synchronize()
        return o, int(lib.csr_launch_count(0))

    _lib.set_option("graph", 0)
    try:
        x0 = torch.zeros((P, plan.ld2n), device=dev)
        ref, n_ref = run(x0)
        ref = {key: t.clone() for key, t in ref.items() if t is not None}
    finally:
        _lib.set_option("graph", 1)
    x = torch.zeros((P, plan.ld2n), device=dev)
    out, counts = None, []
    for call in range(4):      # plain, capture + launch, replay, replay
        out, n = run(x, out)
        counts.append(n)
        assert torch.equal(x, x0), call
        for key, t in ref.items():
            if key == "delta":   # (sums of atomics: the order of the additions is not fixed)
                assert torch.allclose(out[key], t, rtol=1e-4, atol=0), (call, key)
            else:
                assert torch.equal(out[key], t), (call, key)
    assert counts == [n_ref] * 4 and 0 < torch.count_nonzero(x) < 0.2 * x.numel()
    # a replay reads the buffers as they are now
    data.mul_(0.5)
    lam.mul_(0.5)
    nz.mul_(0.5)
    out, _ = run(x, out)
    assert rel(x.cpu().numpy(), 0.5 * x0.cpu().numpy()) < 1e-5 and not torch.equal(x, x0)


@pytest.mark.parametrize("per_pixel", [False, True])
def test_incremental_screening_skips_work_and_keeps_the_bits(per_pixel):
    """incremental screening on a slab large enough for the pair kernels (P > 128): the screening GEMMs issue fewer
    tensor-core blocks than with the switch off, and the solution is the same bit for bit.  Shared weights take the fp8
    cascade, per-pixel weights the one-pass fp16 screening."""
    from csromer_b200 import _device as dv
    from csromer_b200 import _lib
    K, P = 60, 300
    src, par = make_batch(400, P, seed=5, mixed=True, per_pixel_w=per_pixel, flags=0.1 if per_pixel else 0.0)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    plan = dft.device_plan()
    dev = plan.device
    data, _ = dv.pack_planar(src.data, plan.m, plan.ld2m, dev)
    w = dft._channel_tensor(src.w, plan)
    s = torch.as_tensor(np.asarray(src.s), dtype=torch.float32, device=dev)
    k = dv.per_pixel_vector(src.k, P, dev, "k")
    lam0 = 40.0 * np.asarray(src.theo_noise, dtype=np.float64)
    lam = dv.per_pixel_vector(lam0, P, dev, "lambda")
    noise = dv.per_pixel_vector(lam0 / (2.0 * K), P, dev, "noise")
    got = {}
    try:
        for incr in (1, 0):
            _lib.set_option("screen_incr", incr)
            x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=dev)
            out = plan.fista(data, w, s, k, x, lam, noise, K, use_dirty_as_guess=True, want_mma_blocks=True)
            got[incr] = (x.cpu().numpy(), out["mma_blocks"].cpu().numpy().astype(np.float64))
    finally:
        _lib.set_option("screen_incr", 1)
    np.testing.assert_array_equal(got[1][0], got[0][0])
    assert 0 < np.count_nonzero(got[1][0]) < 0.2 * got[1][0].size
    kind = 8 if per_pixel else 9        # fp16 / fp8 screening launches (csr_profile_collect's numbering)
    assert got[0][1][kind] > 0 and got[1][1][kind] < 0.6 * got[0][1][kind], (got[1][1], got[0][1])


def test_per_pixel_spectral_factor_in_the_shared_sampling_path():
    """a [P, m] spectral factor used to be read as a shared [m] vector.  Every line of sight its own
    spectral index, shared weights: GPU loop against the oracle with the same s."""
    from csromer_b200 import _device as dv
    K, P = 25, 37
    src, par = make_batch(200, P, seed=21)
    rng = np.random.RandomState(4)
    alpha = rng.normal(1.0, 0.3, size=P)
    nu_of_l2 = ob.C_LIGHT / np.sqrt(src.lambda2)
    s_pp = (nu_of_l2[:, None] / src.nu_0) ** alpha[None, :]                       # (m, P)
    k_pp = np.sum(np.asarray(src.w)[:, None] / s_pp, axis=0)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    plan = dft.device_plan()
    dev = plan.device
    data, _ = dv.pack_planar(src.data, plan.m, plan.ld2m, dev)
    w = torch.as_tensor(np.asarray(src.w), dtype=torch.float32, device=dev)
    s = torch.as_tensor(s_pp.T.copy(), dtype=torch.float32, device=dev)           # [P, m]
    k = dv.per_pixel_vector(k_pp, P, dev, "k")
    lam0 = 40.0 * float(src.theo_noise)
    lam = dv.per_pixel_vector(lam0, P, dev, "lambda")
    noise = dv.per_pixel_vector(lam0 / (2.0 * K), P, dev, "noise")
    x = torch.zeros((P, plan.ld2n), dtype=torch.float32, device=dev)
    out = plan.fista(data, w, s, k, x, lam, noise, K, use_dirty_as_guess=True)
    op = ob.ExactDFT(src.lambda2, par.phi, src.w, s_pp, k_pp, src.l2_ref)
    d = np.asarray(src.data)
    ref = ob.fista(op, d, ob.complex_to_real(op.backward(d)), lam0, lam0 / (2.0 * K), K)
    got = x[:, :2 * plan.n].double().cpu().numpy().T
    assert rel(got, ref["x"]) < TOL_MODEL
    model = out["model"].double().cpu().numpy()
    assert rel((model[:, :plan.m] + 1j * model[:, plan.m:2 * plan.m]).T, ref["model"]) < TOL_MODEL


def test_sparse_output_equals_dense_output():
    """FISTA(sparse_out=True): the compressed rows hold exactly the non-zeros of the dense solution (NumPy in -> host
    arrays, torch in -> device tensors); a prefetched dataset gives the same answer as a lazily uploaded one."""
    src, par = make_batch(200, 77, seed=13, mixed=True)
    K = 30
    cost, X, l1, lam0, noise = run_fista(src, par, K)
    dense = np.asarray(X.data)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
    guess.data = dft.backward(src.data)
    guess.complex_data_to_real()
    chi2, l1b = cs.Chi2(dft_obj=dft), cs.L1(reg=lam0)
    cost2, Xs = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1b]), fx=chi2, gx=cs.OFunction([l1b]), noise=noise, maxiter=K,
                         verbose=False, sparse_out=True).run()
    rows = Xs.data
    assert isinstance(rows, cs.SparseRows) and isinstance(rows.values, np.ndarray) and len(rows) == dense.shape[0]
    assert rows.shape == dense.shape and rows.nnz == np.count_nonzero(dense) and 0 < rows.nnz < 0.2 * dense.size
    np.testing.assert_array_equal(rows.todense(), dense)
    np.testing.assert_array_equal(cost2, cost)
    idx, val = rows.row(5)
    np.testing.assert_array_equal(idx, np.nonzero(dense[:, 5])[0])
    np.testing.assert_array_equal(val, dense[idx, 5])
    # torch host tensor in (pinned), prefetched on the side stream: device-resident compressed rows out
    data_t = torch.as_tensor(np.asarray(src.data).astype(np.complex64)).pin_memory()
    ds = cs.Dataset(lambda2=src.lambda2, data=data_t, sigma=src.sigma, spectral_idx=1.0).prefetch()
    dft2 = cs.NDFT1D(dataset=ds, parameter=par)
    g2 = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
    g2.data = dft2.backward(ds.data)
    g2.complex_data_to_real()
    chi2c, l1c = cs.Chi2(dft_obj=dft2), cs.L1(reg=lam0)
    _, Xt = cs.FISTA(guess_param=g2, F_obj=cs.OFunction([chi2c, l1c]), fx=chi2c, gx=cs.OFunction([l1c]), noise=noise, maxiter=K,
                     verbose=False, sparse_out=True).run()
    assert dv_is_cuda(Xt.data.values)
    assert rel(Xt.data.todense(), dense) < 1e-6            # (complex64 input instead of complex128: not the same bits)


def dv_is_cuda(t):
    return isinstance(t, torch.Tensor) and t.is_cuda


def test_operand_scales_follow_a_diverging_run():
    """unit step below the stability limit (lambda0 = 5 theo_noise) diverges -- in the
    reference too.  The fp16 operand scales must follow the growth instead of saturating: the GPU run stays finite and
    tracks the float64 oracle while the iterate grows by orders of magnitude."""
    src, par = make_batch(200, 40, seed=3, mixed=True)
    K = 20
    cost, X, l1, lam0, noise = run_fista(src, par, K, lam_mult=5.0)
    op = oracle_operator(src, par)
    start = ob.complex_to_real(op.backward(np.asarray(src.data)))
    ref = ob.fista(op, np.asarray(src.data), start, lam0, noise, K)
    growth = np.abs(ref["x"]).max() / np.abs(start).max()
    assert 1e6 < growth < 1e20, growth     # far beyond the 2^10 headroom of the initial fp16 operand scales
    assert np.isfinite(X.data).all()
    assert rel(X.data, ref["x"]) < 5e-3


def test_fista_on_a_phi_grid_without_mirror_symmetry():
    """the mirror-symmetric screening kernels need phi_j = -phi_{n-j}; any other grid must take the plain kernels (and a
    grid that is symmetric must give the same answer with the shortcut off)"""
    from csromer_b200 import _lib
    src, par = make_batch(200, 140, seed=31, mixed=True)
    par.phi = par.phi + 0.37 * par.cellsize          # shifted grid: no cell mirrors another
    cost, X, l1, lam0, noise = run_fista(src, par, 30)
    ref = oracle_fista(src, par, 30, lam0, noise)
    check_fista(src, X, cost, l1, ref)
    assert np.count_nonzero(X.data) < 0.2 * X.data.size
    src2, par2 = make_batch(200, 140, seed=31, mixed=True)
    res = []
    try:
        for sym in (1, 0):
            _lib.set_option("sym", sym)
            cost2, X2, _, _, _ = run_fista(src2, par2, 30)
            res.append(np.array(X2.data, copy=True))
    finally:
        _lib.set_option("sym", 1)
    np.testing.assert_array_equal(res[0], res[1])


def test_wavelet_forward_k_skipping_is_bit_exact():
    """undecimated basis: the fused synthesis kernel flags the non-zero operand blocks and the forward GEMM skips the
    rest; coefficient-state flags let the update kernel skip state traffic of blocks that are and stay zero -- same
    bits as the dense loops"""
    from csromer_b200 import _lib
    src, par = make_batch(300, 200, seed=77, mixed=True)
    results = []
    try:
        for k_skip, zero_skip in ((1, 1), (0, 1), (0, 0)):
            _lib.set_option("k_skip", k_skip)
            _lib.set_option("zero_skip", zero_skip)
            wav = cs.UndecimatedWavelet(wavelet_name="coif2")
            cost, X, l1, lam0, noise = run_fista(src, par, 25, wav=wav, step=1.0 / 2.8)
            results.append((np.array(X.data, copy=True), np.array(cost, copy=True), np.array(src.model_data, copy=True)))
    finally:
        _lib.set_option("k_skip", 1)
        _lib.set_option("zero_skip", 1)
    assert 0 < np.count_nonzero(results[0][0]) < 0.2 * results[0][0].size
    for other in results[1:]:
        for a, b in zip(results[0], other):
            np.testing.assert_array_equal(a, b)


def test_fista_zero_start_and_step_extension():
    src, par = make_batch(200, 37, seed=5)
    cost, X, l1, lam0, noise = run_fista(src, par, 60, lam_mult=5.0, step=1.0 / 2.8, x0="zeros")
    ref = oracle_fista(src, par, 60, lam0, noise, step=1.0 / 2.8, x0="zeros")
    check_fista(src, X, cost, l1, ref)


@pytest.mark.parametrize("kind,name,kw", [
    ("swt", "coif2", {}), ("swt", "sym2", dict(append_signal=True)), ("swt", "db4", dict(norm=True)),
    ("dwt", "db2", {}), ("dwt", "coif2", dict(append_signal=True)),
])
def test_fista_wavelet_bases(kind, name, kw):
    # append_signal doubles the synthesis operator's norm ([I, W]): the unit step diverges there (the oracle too),
    # so those cases run with the `step` extension on both sides
    step = 0.3 if kw.get("append_signal") else 1.0
    src, par = make_batch(200, 40, seed=21, mixed=True)
    if kind == "swt":
        wav = cs.UndecimatedWavelet(wavelet_name=name, **kw)
    else:
        wav = cs.DiscreteWavelet(wavelet_name=name, mode="periodization", **kw)
    owav = ob.WaveletDict(name, kind=kind, norm=kw.get("norm", False), append_signal=kw.get("append_signal", False))
    cost, X, l1, lam0, noise = run_fista(src, par, 30, wav=wav, step=step)
    ref = oracle_fista(src, par, 30, lam0, noise, owav=owav, step=step)
    check_fista(src, X, cost, l1, ref)
    # synthesis of the solution, as the README does after run() (README.md:247-248)
    sig = wav.reconstruct(X.data)
    assert rel(sig, owav.reconstruct(ref["x"])) < TOL_MODEL


def test_generic_python_loop_equals_fused_loop():
    """FISTA(fused=False) drives Chi2 / L1 / OFunction method by method (the reference's loop, every transform on
    the GPU); it must agree with the fused device loop and with the oracle."""
    src, par = make_batch(200, 1, seed=3)
    src1 = cs.Dataset(lambda2=src.lambda2, data=np.asarray(src.data)[:, 0], sigma=src.sigma, spectral_idx=1.0)
    c_f, X_f, l1_f, lam0, noise = run_fista(src1, par, 15, fused=True)
    model_fused = np.array(src1.model_data)
    c_g, X_g, l1_g, _, _ = run_fista(src1, par, 15, fused=False)
    assert X_g.data.shape == X_f.data.shape == (2 * len(par.phi),)
    assert rel(X_g.data, X_f.data) < TOL_MODEL and rel(src1.model_data, model_fused) < TOL_MODEL
    assert c_g == pytest.approx(c_f, rel=1e-4) and l1_g.reg == pytest.approx(l1_f.reg, rel=1e-12)
    wav = cs.UndecimatedWavelet(wavelet_name="db2", wavelet_level=3)
    c_f, X_f, _, _, _ = run_fista(src1, par, 10, wav=wav, fused=True)
    c_g, X_g, _, _, _ = run_fista(src1, par, 10, wav=wav, fused=False)
    assert rel(X_g.data, X_f.data) < TOL_MODEL and c_g == pytest.approx(c_f, rel=1e-4)


def test_fista_iteration_count_semantics():
    """maxiter=None -> floor(lambda/noise) iterations per line of sight (fista.py:59-66); noise >= lambda returns
    the guess untouched with cost 0 (fista.py:74-77)."""
    src, par = make_batch(200, 6, seed=9)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    lam0 = 40.0 * np.asarray(src.theo_noise) * np.array([1.0, 1.0, 0.5, 0.5, 0.25, 2.0])
    noise = 40.0 * np.asarray(src.theo_noise) / 16.3  # -> 16, 16, 8, 8, 4, 32 iterations
    guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
    guess.data = dft.backward(src.data)
    guess.complex_data_to_real()
    chi2, l1 = cs.Chi2(dft_obj=dft), cs.L1(reg=lam0.copy())
    cost, X = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                       verbose=False).run()
    op = oracle_operator(src, par)
    for p in range(6):
        opp = ob.ExactDFT(src.lambda2, par.phi, src.w, src.s, src.k)
        rp = ob.fista(opp, np.asarray(src.data)[:, p], np.asarray(guess.data, dtype=np.float64)[:, p], lam0[p], noise)
        assert rp["iters"] == int(np.floor(lam0[p] / noise))
        assert rel(X.data[:, p], rp["x"]) < TOL_MODEL
        assert l1.reg[p] == pytest.approx(rp["lam_final"], rel=1e-9)
    l1b = cs.L1(reg=1e-6)
    before = np.array(guess.data)
    cost, X = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1b]), fx=chi2, gx=cs.OFunction([l1b]), noise=1e-3,
                       maxiter=5, verbose=False).run()
    assert cost == 0.0 and np.array_equal(X.data, before) and l1b.reg == 1e-6


def make_ragged_stack(P, nch, seed, keep=0.8):
    """P lines of sight that each keep their own random subset of the channels (what a flagger with
    delete_channels=True leaves behind), one with a non-zero l2_ref, on the shared phi grid of the full band"""
    rng = np.random.RandomState(seed)
    nu = np.linspace(1.008e9, 2.031e9, nch)
    full = cs.Dataset(nu=nu, sigma=np.ones(nch), spectral_idx=1.0)
    par = cs.Parameter()
    par.calculate_cellsize(dataset=full, oversampling=8, verbose=False)
    l2 = np.asarray(full.lambda2)
    members = []
    for p in range(P):
        idx = np.sort(rng.choice(nch, size=max(8, int(keep * nch) - 3 * p), replace=False)) if p else np.arange(nch)
        lam = l2[idx]
        sigma = 0.0007 * rng.uniform(0.5, 2.0, size=idx.size)
        phi0, amp = rng.uniform(-400, 400), 0.0035 * rng.uniform(0.5, 1.5)
        ds = cs.Dataset(lambda2=lam, sigma=sigma, spectral_idx=1.0, l2_ref=0.03 if p == 1 else None)
        clean = amp * ds.s * np.exp(2j * phi0 * (lam - ds.l2_ref))
        ds.data = (clean + rng.normal(0, 0.0007, idx.size) + 1j * rng.normal(0, 0.0007, idx.size)).astype(np.complex64)
        if p == 2:
            ds.w[5:15] = 0.0  # in-place flagging on top (k stays, like the reference)
        members.append(ds)
    return members, par


def test_per_pixel_sampling_transforms_and_fista():
    """K4: every line of sight has its own lambda^2 list.  Transforms and the FISTA loop against the oracle run
    line of sight by line of sight (the README's per-pixel loop)."""
    P, nch, K = 7, 160, 30
    members, par = make_ragged_stack(P, nch, seed=3)
    op = cs.PerPixelDFT1D(dataset=members, parameter=par)
    n = len(par.phi)
    stack = op.dataset
    F = op.backward(stack.data)
    R = op.RMTF()
    rng = np.random.RandomState(9)
    x = (rng.normal(size=(n, P)) + 1j * rng.normal(size=(n, P))) * (rng.uniform(size=(n, P)) < 0.03)
    fw, fn = op.forward(x), op.forward_normalized(x)
    assert F.shape == (n, P) and fw.shape == (nch, P)
    for p, ds in enumerate(members):
        o = ob.ExactDFT(ds.lambda2, par.phi, ds.w, ds.s, ds.k, ds.l2_ref)
        assert rel(F[:, p], o.backward(np.asarray(ds.data))) < TOL_TRANSFORM
        assert rel(R[:, p], o.rmtf()) < TOL_TRANSFORM
        assert rel(fw[:ds.m, p], o.forward(x[:, p])) < TOL_TRANSFORM
        good = np.asarray(ds.w) > 0
        assert rel(fn[:ds.m, p][good], o.forward_normalized(x[:, p])[good]) < TOL_TRANSFORM
        assert np.all(fw[ds.m:, p] == 0) and np.all(fn[:ds.m, p][~good] == 0)
    for wav_name in (None, "coif2"):
        for ds in members:
            ds.model_data = np.zeros_like(ds.data)
        lam0 = 40.0 * stack.theo_noise
        noise = lam0 / (2.0 * K)
        wav = cs.UndecimatedWavelet(wavelet_name=wav_name) if wav_name else None
        guess = cs.Parameter(phi=par.phi, cellsize=par.cellsize)
        guess.data = F
        guess.complex_data_to_real()
        if wav is not None:
            guess.data = wav.decompose(guess.data)
        chi2 = cs.Chi2(dft_obj=op, wavelet=wav)
        l1 = cs.L1(reg=lam0)
        step = 1.0 if wav is None else 1.0 / 2.8
        cost, X = cs.FISTA(guess_param=guess, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                           maxiter=K, verbose=False, step=step).run()
        for p, ds in enumerate(members):
            o = ob.ExactDFT(ds.lambda2, par.phi, ds.w, ds.s, ds.k, ds.l2_ref)
            start = ob.complex_to_real(o.backward(np.asarray(ds.data)))
            owav = None
            if wav is not None:
                owav = ob.WaveletDict("coif2", kind="swt")
                start = owav.decompose(start)
            ref = ob.fista(o, np.asarray(ds.data), start, lam0[p], noise[p], K, wavelet=owav, step=step)
            assert np.isfinite(ref["x"]).all()
            good = np.asarray(ds.w) > 0
            assert rel(np.asarray(X.data)[:, p], ref["x"]) < TOL_MODEL
            assert rel(np.asarray(ds.model_data)[good], ref["model"][good]) < TOL_MODEL
            assert rel(np.asarray(cost)[p], ref["cost"]) < TOL_MODEL


def test_nufft_kernels_against_the_exact_sums():
    """K4n (csrc/nufft.cu): the shared-memory NUFFT of the per-line-of-sight transforms against the direct sums of
    csrc/nudft.cu on ragged channel lists -- single transforms within the stated 1e-5 (observed ~1e-6), FISTA within the
    model tolerance, and <forward x, b> = <x, adjoint b> (type 1 is the exact adjoint of type 2)."""
    from csromer_b200 import _device as dv
    from csromer_b200 import _lib
    P, nch = 23, 300
    members, par = make_ragged_stack(P, nch, seed=41)
    stack = cs.DatasetStack(members)
    st = dv.DeviceStack(stack.lambda2_minus_ref.T, stack.m_count, par.phi)
    g = torch.Generator(device=st.device).manual_seed(4)
    x = torch.randn((P, st.ld2n), generator=g, device=st.device)
    x[:, 2 * st.n:] = 0
    b = torch.randn((P, st.ld2m), generator=g, device=st.device)
    valid = (torch.arange(st.m, device=st.device)[None, :] < st.m_count[:, None]).float()
    b[:, :st.m] *= valid
    b[:, st.m:2 * st.m] *= valid
    b[:, 2 * st.m:] = 0
    cs_t = torch.rand((P, st.m), generator=g, device=st.device) + 0.5
    rs_t = torch.rand((P,), generator=g, device=st.device) + 0.5
    out = {}
    try:
        for nufft in (1, 0):
            _lib.set_option("nufft", nufft)
            out[nufft] = (st.forward(x, cs_t, rs_t), st.adjoint(b, cs_t, rs_t), st.forward(x), st.adjoint(b))
    finally:
        _lib.set_option("nufft", 1)
    for fast, exact in zip(out[1], out[0]):
        assert float((fast - exact).abs().max() / exact.abs().max()) < TOL_TRANSFORM
    fwd, adj = out[1][2], out[1][3]
    lhs = float((fwd.double() * b.double()).sum())
    rhs = float((x.double() * adj.double()).sum())
    assert abs(lhs - rhs) < 2e-5 * max(abs(lhs), abs(rhs), float(fwd.double().norm() * b.double().norm()))


def test_galactic_rm_derotation_on_device():
    from csromer_b200 import _device as dv
    src, par = make_batch(200, 90, seed=13)
    plan = cs.NDFT1D(dataset=src, parameter=par).device_plan()
    rng = np.random.RandomState(0)
    phi_gal = rng.normal(0, 30, size=90)
    d, pix = dv.pack_planar(src.data, plan.m, plan.ld2m, plan.device)
    plan.derotate(d, torch.as_tensor(phi_gal, dtype=torch.float32, device=plan.device))
    got = dv.unpack_complex(d, plan.m, pix)
    want = ob.subtract_galactic_rm(np.asarray(src.data), src.lambda2, phi_gal.astype(np.float32).astype(np.float64))
    assert rel(got, want) < 1e-6


def test_torch_tensors_stay_on_device():
    src, par = make_batch(200, 20, seed=17)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    d = torch.as_tensor(np.asarray(src.data)).cuda()
    F = dft.backward(d)
    assert isinstance(F, torch.Tensor) and F.is_cuda and F.dtype == torch.complex64
    assert rel(F.cpu().numpy(), oracle_operator(src, par).backward(np.asarray(src.data))) < TOL_TRANSFORM


# ---------------------------------------------------------------------------------------------------------
# wavelet dictionaries
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kind,N,level,norm,app", [
    ("db1", "swt", 64, None, False, False), ("coif2", "swt", 3136, None, False, False),
    ("sym2", "swt", 3136, 3, False, True), ("db4", "swt", 3136, None, True, False),
    ("bior2.2", "swt", 1024, 4, False, False), ("IUWT", "swt", 512, 3, False, False),
    ("db2", "dwt", 3136, None, False, False), ("coif2", "dwt", 15936, None, False, True),
    ("db1", "dwt", 250, None, False, False), ("sym4", "dwt", 1001, 3, False, False),
])
def test_wavelet_transforms(name, kind, N, level, norm, app):
    rng = np.random.RandomState(2)
    x = rng.normal(size=(N, 9)).astype(np.float32)
    if kind == "swt":
        w = cs.UndecimatedWavelet(wavelet_name=name, wavelet_level=level, append_signal=app, norm=norm)
    else:
        w = cs.DiscreteWavelet(wavelet_name=name, wavelet_level=level, mode="periodization", append_signal=app)
    o = ob.WaveletDict(name, kind=kind, level=level, norm=norm, append_signal=app)
    c, co = w.decompose(x), o.decompose(x.astype(np.float64))
    assert c.shape == co.shape and w.ncoeffs == co.shape[0] and w.n == N
    assert rel(c, co) < 2e-6
    r = w.reconstruct(c)
    # (PyWavelets returns N+1 samples for an odd-length periodization round trip; this package returns the N it was
    # given -- lengths on the hot path are always even, 2n)
    assert rel(r, o.reconstruct(co)[:N]) < 2e-6
    if name != "IUWT":  # the reference's IUWT bank is not perfectly reconstructing
        assert rel(r, (2 if app else 1) * x) < 2e-6
    # single line of sight keeps 1-D shapes
    c1 = w.decompose(x[:, 0])
    assert c1.shape == (co.shape[0],) and rel(c1, co[:, 0]) < 2e-6
    with pytest.raises(ValueError):
        w.reconstruct(c[:-1])


@pytest.mark.parametrize("mode", ["zero", "constant", "symmetric", "reflect", "periodic", "smooth", "antisymmetric", "antireflect"])
def test_decimated_wavelet_signal_extension_modes(mode):
    """DiscreteWavelet(mode=...) hands the mode to pywt.wavedec / waverec (discrete.py:37-39,80-89): bands of
    (len + F - 1) // 2 samples, the extension-independent part of the inverse; checked against the oracle's restatement
    (pinned on the PyWavelets documentation value in tests/test_oracle_wavelets.py)."""
    rng = np.random.RandomState(31)
    for name, N, level, app in (("db2", 3136, None, False), ("coif2", 1000, 4, True), ("db1", 250, None, False), ("sym4", 37, 2, False),
                                ("db4", 17, 1, False)):
        x = rng.normal(size=(N, 7)).astype(np.float32)
        w = cs.DiscreteWavelet(wavelet_name=name, wavelet_level=level, mode=mode, append_signal=app)
        o = ob.WaveletDict(name, kind="dwt", level=level, append_signal=app, mode=mode)
        c, co = w.decompose(x), o.decompose(x.astype(np.float64))
        assert c.shape == co.shape and w.ncoeffs == co.shape[0] and w.n == N
        sizes = [sh[0] for sh in w.coeff_shapes]
        assert sizes == o.shapes and sizes[-1] == w.calculate_ncoeffs(x[:, 0])
        tol = 2e-5 if mode in ("smooth", "antireflect") else 3e-6  # extrapolated edges are large next to the signal
        assert rel(c, co) < tol
        r = w.reconstruct(c)
        assert r.shape == x.shape and rel(r, o.reconstruct(co)) < tol and rel(r, (2 if app else 1) * x) < tol
    # the complex variants transform real and imaginary parts separately (discrete.py:46-68)
    z = (rng.normal(size=200) + 1j * rng.normal(size=200)).astype(np.complex64)
    w = cs.DiscreteWavelet(wavelet_name="db3", mode=mode)
    cz = w.decompose_complex(z)
    o = ob.WaveletDict("db3", kind="dwt", mode=mode)
    assert rel(cz.real, o.decompose(z.real.astype(np.float64))) < 2e-5 and rel(w.reconstruct_complex(cz), z) < 2e-5


@pytest.mark.parametrize("mode", ["periodization", "symmetric", "zero", "smooth"])
def test_decimated_transform_fused_kernels_are_bit_exact(mode):
    """all levels of the decimated transform in one launch (the running approximation stays in shared memory): same bits as the
    level-by-level kernels, one launch per transform (two with append_signal's nothing extra: the copy is fused too)"""
    from csromer_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(41)
    for name, N, level, app in (("db2", 15936, None, False), ("coif2", 3136, None, True), ("sym4", 1001, 3, False), ("db1", 250, None, False)):
        x = rng.normal(size=(N, 5)).astype(np.float32)
        out = []
        for fused in (1, 0):
            _lib.set_option("fused_wavelet", fused)
            try:
                w = cs.DiscreteWavelet(wavelet_name=name, wavelet_level=level, mode=mode, append_signal=app)
                lib.csr_launch_count(1)
                c = w.decompose(x)
                n_dec = int(lib.csr_launch_count(1))
                r = w.reconstruct(c)
                n_rec = int(lib.csr_launch_count(0))
            finally:
                _lib.set_option("fused_wavelet", 1)
            out.append((c, r, n_dec, n_rec, len(w.coeff_shapes) - 1))
        (c1, r1, d1, k1, L), (c0, r0, d0, k0, _) = out
        np.testing.assert_array_equal(c1, c0)
        np.testing.assert_array_equal(r1, r0)
        # (+ the two transposes between the Python-side layout and the planar rows)
        assert d0 - d1 == L - 1 + (1 if app else 0) and k0 - k1 == L - 1, (d1, d0, k1, k0, L)


def test_fista_in_a_symmetric_extension_wavelet_basis():
    src, par = make_batch(200, 24, seed=22, mixed=True)
    wav = cs.DiscreteWavelet(wavelet_name="db2", mode="symmetric")
    owav = ob.WaveletDict("db2", kind="dwt", mode="symmetric")
    cost, X, l1, lam0, noise = run_fista(src, par, 30, wav=wav)
    ref = oracle_fista(src, par, 30, lam0, noise, owav=owav)
    check_fista(src, X, cost, l1, ref)
    assert rel(wav.reconstruct(X.data), owav.reconstruct(ref["x"])) < TOL_MODEL


@pytest.mark.parametrize("name,N,level", [("db1", 4096, None), ("sym2", 3136, None), ("coif1", 3136, 4), ("db4", 15936, None),
                                          ("coif2", 15936, None), ("coif2", 63936, None), ("coif2", 1032, 3)])
def test_wavelet_vec4_kernels_are_bit_exact(name, N, level):
    """the fused undecimated kernels with four samples per thread (16-byte shared / global accesses) accumulate every
    output tap by tap like the one-sample kernels: same bits out, for the transforms and for a FISTA run"""
    from csromer_b200 import _lib
    rng = np.random.RandomState(11)
    x = rng.normal(size=(N, 5)).astype(np.float32)
    x[N // 3: N // 3 + N // 4, 1] = 0.0
    out = []
    try:
        for vec4 in (1, 0):
            _lib.set_option("wavelet_vec4", vec4)
            w = cs.UndecimatedWavelet(wavelet_name=name, wavelet_level=level)
            c = w.decompose(x)
            c_sparse = np.where(np.abs(c) > 2.0, c, 0).astype(np.float32)  # mostly-zero band tiles (the skip paths)
            out.append((np.array(c, copy=True), np.array(w.reconstruct(c), copy=True), np.array(w.reconstruct(c_sparse), copy=True)))
    finally:
        _lib.set_option("wavelet_vec4", 1)
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)
    o = ob.WaveletDict(name, kind="swt", level=level)
    assert rel(out[0][0], o.decompose(x.astype(np.float64))) < 2e-6


def test_wavelet_vec4_fista_is_bit_exact():
    from csromer_b200 import _lib
    src, par = make_batch(300, 70, seed=78, mixed=True)
    results = []
    try:
        for vec4 in (1, 0):
            _lib.set_option("wavelet_vec4", vec4)
            wav = cs.UndecimatedWavelet(wavelet_name="coif2")
            cost, X, l1, lam0, noise = run_fista(src, par, 25, wav=wav, step=1.0 / 2.8)
            results.append((np.array(X.data, copy=True), np.array(cost, copy=True), np.array(src.model_data, copy=True)))
    finally:
        _lib.set_option("wavelet_vec4", 1)
    assert 0 < np.count_nonzero(results[0][0]) < 0.2 * results[0][0].size
    for a, b in zip(results[0], results[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("nch,P,K,name,kw", [(300, 70, 25, "coif2", dict(mixed=True)), (1000, 261, 12, "coif2", dict(mixed=True)),
                                              (203, 300, 30, "sym2", dict(per_pixel_w=True, flags=0.1)), (400, 129, 40, "db4", {})])
def test_wavelet_domain_screening_is_bit_exact(nch, P, K, name, kw):
    """undecimated basis: away from the coefficients with non-zero state the gradient is computed with ONE fp16 product
    and a rigorous bound proves that the coefficients there stay zero; blocks that fail are redone with the exact
    gradient.  Same bits as the dense three-product loop, and parity with the oracle."""
    from csromer_b200 import _lib
    src, par = make_batch(nch, P, seed=79, **kw)
    results = []
    try:
        for screen, pair in ((1, 1), (1, 0), (0, 1)):  # (pair: the one-product gradient runs on cta_group::2 CTA pairs)
            _lib.set_option("wavelet_screen", screen)
            _lib.set_option("pair", pair)
            wav = cs.UndecimatedWavelet(wavelet_name=name)
            cost, X, l1, lam0, noise = run_fista(src, par, K, wav=wav, step=1.0 / 2.8)
            results.append((np.array(X.data, copy=True), np.array(cost, copy=True), np.array(src.model_data, copy=True)))
    finally:
        _lib.set_option("wavelet_screen", 1)
        _lib.set_option("pair", 1)
    assert 0 < np.count_nonzero(results[0][0]) < 0.3 * results[0][0].size
    for other in results[1:]:
        for a, b in zip(results[0], other):
            np.testing.assert_array_equal(a, b)
    if nch <= 400:
        owav = ob.WaveletDict(name, kind="swt")
        ref = oracle_fista(src, par, K, lam0, noise, owav=owav, step=1.0 / 2.8)
        check_fista(src, X, cost, l1, ref)


def test_wavelet_complex_variants_and_errors():
    rng = np.random.RandomState(4)
    z = (rng.normal(size=256) + 1j * rng.normal(size=256)).astype(np.complex64)
    w = cs.UndecimatedWavelet(wavelet_name="db2", wavelet_level=4)
    c = w.decompose_complex(z)
    assert c.shape == (5 * 256,) and rel(w.reconstruct_complex(c), z) < 2e-6
    with pytest.raises(ValueError):
        cs.UndecimatedWavelet(wavelet_name="db2", wavelet_level=9).decompose(np.zeros(256, dtype=np.float32))
    with pytest.raises(RuntimeError):
        cs.DiscreteWavelet(wavelet_name="db2", mode="periodization").reconstruct(np.zeros(8, dtype=np.float32))


# ---------------------------------------------------------------------------------------------------------
# size-independent properties at a bench-like size (C3 sampling: 1000 channels, n = 7968)
# ---------------------------------------------------------------------------------------------------------
def test_full_size_linearity_and_adjointness():
    from csromer_b200 import _device as dv
    nu = np.linspace(1.008e9, 2.031e9, 1000)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200.0, spectral_idx=1.0)
    src.simulate()
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    plan = dv.DevicePlan.get(src.lambda2, 0.0, par.phi)
    assert (plan.m, plan.n) == (1000, 7968)
    P = 1500
    g = torch.Generator(device=plan.device).manual_seed(1234)
    x = torch.randn((P, plan.ld2n), generator=g, device=plan.device)
    y = torch.randn((P, plan.ld2m), generator=g, device=plan.device)
    Ex = plan.forward(x)
    Ehy = plan.adjoint(y)
    # <E x, y> = <x, E^H y> for every line of sight (real inner product of the planar vectors)
    lhs = (Ex.double() * y.double()).sum(dim=1)
    rhs = (x.double() * Ehy.double()).sum(dim=1)
    scale = (Ex.double().norm(dim=1) * y.double().norm(dim=1))
    assert float(((lhs - rhs).abs() / scale).max()) < 1e-6
    # linearity
    x2 = torch.randn((P, plan.ld2n), generator=g, device=plan.device)
    comb = plan.forward(2.0 * x - 0.5 * x2)
    want = 2.0 * Ex - 0.5 * plan.forward(x2)
    assert float((comb - want).abs().max() / want.abs().max()) < 1e-5
    # (1/n) E^H E has trace m/n * n = m on average: diag of the normal operator equals m/n
    e = torch.zeros((8, plan.ld2n), device=plan.device)
    idx = torch.tensor([0, 1, 100, 3984, 5000, 7967, 7968, 15935], device=plan.device)
    e[torch.arange(8, device=plan.device), idx] = 1.0
    diag = plan.adjoint(plan.forward(e))[torch.arange(8, device=plan.device), idx] / plan.n
    assert float((diag - plan.m / plan.n).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# the BASELINE.json configurations beyond C1/C3 (reduced pixel counts; full samplings)
# ---------------------------------------------------------------------------------------------------------
def test_config2_single_los_wavelet_flagged_nufft():
    """C2: one line of sight, thin + thick mixed source, undecimated coif2 basis, 20 % of the channels flagged as
    w = 0 in RFI-like chunks, NUFFT1D API (exact operator)."""
    rng = np.random.RandomState(42)
    nu = np.linspace(1.008e9, 2.031e9, 1000)
    thin = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    thick = cs.FaradayThickSource(nu=nu, s_nu=0.0035, phi_fg=140, phi_center=200, spectral_idx=1.0)
    thin.simulate()
    thick.simulate()
    src = thin + thick
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    w = src.w.copy()
    for start in rng.randint(0, 950, size=6):
        w[start:start + 34] = 0.0
    src.w = w
    assert 0.1 < np.mean(src.w == 0) < 0.25
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    nufft = cs.NUFFT1D(dataset=src, parameter=par, solve=True)
    wav = cs.UndecimatedWavelet(wavelet_name="coif2")
    K = 40
    lam0 = 40.0 * src.theo_noise
    noise = lam0 / (2.0 * K)
    par.data = dft.backward(src.data)
    par.complex_data_to_real()
    par.data = wav.decompose(par.data)
    chi2 = cs.Chi2(dft_obj=nufft, wavelet=wav)
    l1 = cs.L1(reg=lam0)
    cost, X = cs.FISTA(guess_param=par, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=noise,
                       maxiter=K, verbose=False).run()
    op = oracle_operator(src, par)
    owav = ob.WaveletDict("coif2", kind="swt")
    ref = ob.fista(op, np.asarray(src.data), owav.decompose(ob.complex_to_real(op.backward(np.asarray(src.data)))), lam0,
                   noise, K, wavelet=owav)
    assert X.data.shape == (7 * 2 * 7968,)
    check_fista(src, X, cost, l1, ref)
    X.data = wav.reconstruct(X.data)
    X.real_data_to_complex()
    assert rel(X.data, ob.real_to_complex(owav.reconstruct(ref["x"]))) < TOL_MODEL


def test_config4_sampling_per_pixel_weights_flags_galactic_rm():
    """C4 sampling (2000 channels, n = 15968): per-pixel sigma, 10-20 % w = 0 flags per pixel, Galactic RM removed
    per pixel with Dataset.subtract_galacticrm semantics; reduced to 48 lines of sight."""
    rng = np.random.RandomState(1235)
    P, K = 48, 12
    nu = np.linspace(1.008e9, 2.031e9, 2000)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035 * rng.uniform(0.5, 1.5, size=P), phi_gal=rng.uniform(-500, 500, size=P),
                               spectral_idx=1.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    src.subtract_galacticrm(rng.normal(0, 30, size=P))
    src.sigma = 0.0007 * rng.uniform(0.5, 2.0, size=(2000, P))
    w = src.w.copy()
    w[rng.uniform(size=w.shape) < rng.uniform(0.1, 0.2, size=P)] = 0.0
    src.w = w
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    assert len(par.phi) == 15968
    cost, X, l1, lam0, noise = run_fista(src, par, K)
    ref = oracle_fista(src, par, K, lam0, noise)
    check_fista(src, X, cost, l1, ref)


def test_config5_sampling_transforms_and_properties():
    """C5 sampling (4000 channels, MeerKAT band, n = 31968): transforms against the oracle for a few lines of
    sight, adjointness at 300, and a short SWT-coif2 FISTA run against the oracle."""
    from csromer_b200 import _device as dv
    rng = np.random.RandomState(1236)
    P = 6
    nu = np.linspace(0.9e9, 1.67e9, 4000)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035 * rng.uniform(0.5, 1.5, size=P), phi_gal=rng.uniform(-500, 500, size=P),
                               spectral_idx=1.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    assert len(par.phi) == 31968
    dft = cs.NDFT1D(dataset=src, parameter=par)
    op = oracle_operator(src, par)
    data = np.asarray(src.data)
    assert rel(dft.backward(data), op.backward(data)) < TOL_TRANSFORM
    x = (rng.normal(size=(31968, P)) + 1j * rng.normal(size=(31968, P))) * (rng.uniform(size=(31968, P)) < 0.01)
    assert rel(dft.forward_normalized(x), op.forward_normalized(x)) < TOL_TRANSFORM
    wav = cs.UndecimatedWavelet(wavelet_name="coif2")
    owav = ob.WaveletDict("coif2", kind="swt")
    K = 6
    cost, X, l1, lam0, noise = run_fista(src, par, K, wav=wav)
    ref = oracle_fista(src, par, K, lam0, noise, owav=owav)
    assert X.data.shape == (7 * 2 * 31968, P)
    check_fista(src, X, cost, l1, ref)
    plan = dv.DevicePlan.get(src.lambda2, 0.0, par.phi)
    g = torch.Generator(device=plan.device).manual_seed(5)
    xs = torch.randn((300, plan.ld2n), generator=g, device=plan.device)
    ys = torch.randn((300, plan.ld2m), generator=g, device=plan.device)
    lhs = (plan.forward(xs).double() * ys.double()).sum(dim=1)
    rhs = (xs.double() * plan.adjoint(ys).double()).sum(dim=1)
    scale = plan.forward(xs).double().norm(dim=1) * ys.double().norm(dim=1)
    assert float(((lhs - rhs).abs() / scale).max()) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# "next" rows: restore beam, FDF-edge noise, peak statistics and the batched cube recipe (README.md:282-377)
# ---------------------------------------------------------------------------------------------------------
def test_parameter_convolve_and_peak_statistics(capsys):
    from csromer_b200 import _device as dv
    src, par = make_batch(200, 19, seed=31, mixed=True)
    dft = cs.NDFT1D(dataset=src, parameter=par)
    F = dft.backward(src.data)
    par.data = F
    got = par.convolve()
    assert "Convolving with Gaussian kernel" in capsys.readouterr().out
    beam = ob.gaussian_beam(par.rmtf_fwhm, par.cellsize)
    assert beam.shape[0] % 2 == 1 and abs(beam.sum() - 1) < 1e-12
    want = np.stack([ob.convolve_beam(F[:, p].astype(np.complex128), beam) for p in range(19)], axis=1)
    assert got.shape == F.shape and rel(got, want) < 2e-6
    one = par.convolve(x=F[:, 3], verbose=False)
    assert one.shape == (len(par.phi),) and rel(one, want[:, 3]) < 2e-6
    plan = dft.device_plan()
    planar, _ = dv.pack_planar(F, plan.n, plan.ld2n, plan.device)
    idx, stats = dv.peak_stats(planar, plan.n)
    idx, stats = idx.cpu().numpy(), stats.cpu().numpy()
    for p in range(19):
        assert idx[p] == np.argmax(np.abs(F[:, p]))
        phi_pk, pk = ob.peak_parabola(F[:, p].astype(np.complex128), par.cellsize)
        assert (idx[p] + stats[p, 1] - plan.n / 2) * par.cellsize == pytest.approx(phi_pk, abs=1e-3 * par.cellsize)
        assert stats[p, 2] == pytest.approx(pk, rel=1e-5)
    edge = np.abs(par.phi) > par.max_faraday_depth / 1.5
    noise = dv.edge_noise(planar, plan.n, torch.as_tensor(edge.astype(np.uint8)).cuda(), 1.0).cpu().numpy()
    for p in range(19):
        assert noise[p] == pytest.approx(ob.edge_noise(F[:, p].astype(np.complex128), par.phi, par.max_faraday_depth), rel=1e-4)


@pytest.mark.parametrize("use_wavelet", [True, False])
def test_cube_recipe_matches_oracle(use_wavelet):
    from csromer_b200.cube import estimate_lipschitz, reconstruct_cube
    rng = np.random.RandomState(77)
    nu = np.linspace(1.008e9, 2.031e9, 200)
    Y, X = 3, 4
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035 * rng.uniform(0.5, 1.5, size=(Y, X)),
                               phi_gal=rng.uniform(-300, 300, size=(Y, X)), spectral_idx=0.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    cube = np.asarray(src.data)
    sigma = np.full(200, 0.2 * 0.0035)
    F, info = reconstruct_cube(cube, nu, sigma=sigma, use_wavelet=use_wavelet, slab=7)
    assert F.shape == (4, len(info["phi"]), Y, X) and F.dtype == np.complex64
    assert 2.0 < 1.0 / info["step"] < 3.5  # lambda_max((1/n) E^H E), 2.74 at the 1000-channel JVLA sampling
    Fo, io = ob.cube_recipe(cube.reshape(200, -1), nu, sigma, wavelet="coif2" if use_wavelet else None, step=info["step"])
    assert np.isfinite(Fo).all()
    Fo = Fo.reshape(F.shape)
    for plane, tol in ((0, TOL_TRANSFORM), (1, TOL_MODEL), (2, TOL_MODEL), (3, TOL_MODEL)):
        assert rel(F[plane], Fo[plane]) < tol, plane
    np.testing.assert_allclose(info["noise"].reshape(-1), io["noise"], rtol=1e-4)
    np.testing.assert_allclose(info["rm_restored"].reshape(-1), io["rm_restored"], atol=1e-9)
    np.testing.assert_allclose(info["peak_restored"].reshape(-1), io["peak_restored"], rtol=1e-4)
    assert info["iterations"] == int(np.floor(np.sqrt(400 + np.sqrt(800))))


def test_cube_recipe_with_blanked_samples():
    """a NaN-blanked pixel must not poison its slab.  One dead pixel (all NaN), one pixel with a few blanked
    channels, one with an inf: the dead pixel comes back as NaN planes, the partly blanked ones as the reconstruction
    with those channels flagged (w = 0) -- checked against the oracle run with those weights -- and every other
    pixel stays within tolerance of the clean run."""
    from csromer_b200.cube import reconstruct_cube
    rng = np.random.RandomState(78)
    nu = np.linspace(1.008e9, 2.031e9, 200)
    Y, X = 2, 4
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035 * rng.uniform(0.5, 1.5, size=(Y, X)),
                               phi_gal=rng.uniform(-300, 300, size=(Y, X)), spectral_idx=0.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=rng)
    clean = np.asarray(src.data).astype(np.complex128)
    sigma = np.full(200, 0.2 * 0.0035)
    cube = clean.copy()
    cube[:, 0, 1] = np.nan                      # dead pixel
    cube[[5, 17, 18, 150], 1, 2] = np.nan       # four blanked channels
    cube[60, 0, 3] = np.inf                     # one inf
    F0, info0 = reconstruct_cube(clean, nu, sigma=sigma, use_wavelet=False, slab=8)
    F, info = reconstruct_cube(cube, nu, sigma=sigma, use_wavelet=False, slab=8, step=info0["step"])
    assert info["blanked_samples"] == 200 + 4 + 1 and info["dead_pixels"].sum() == 1 and info["dead_pixels"][0, 1]
    assert np.isnan(F[:, :, 0, 1]).all()
    good = np.ones((Y, X), dtype=bool)
    good[0, 1] = False
    assert np.isfinite(F[:, :, good]).all()
    untouched = good.copy()
    untouched[1, 2] = untouched[0, 3] = False
    assert rel(F[:, :, untouched], F0[:, :, untouched]) < TOL_MODEL
    # the partly blanked pixel against the oracle with those channels at weight 0 (same phi grid as the cube run)
    l2 = ob.lambda2_from_nu(nu)
    order = np.argsort(-nu)                     # lambda2 is stored ascending; data stays in caller order
    sc = ob.dataset_scalars(l2, sigma, 0.0)
    wp = sc["w"].copy()
    d = cube[:, 1, 2].copy()
    bad = ~np.isfinite(d)
    d[bad] = 0.0
    wp[bad] = 0.0
    op = ob.ExactDFT(l2, info["phi"], wp, sc["s"])
    dirty = op.backward(d)
    assert rel(F[0, :, 1, 2], dirty) < TOL_TRANSFORM
    grid = ob.phi_grid(l2, sc["w"], 8.0)
    noise = ob.edge_noise(dirty, info["phi"], grid["max_faraday_depth"], 1.0)
    assert info["noise"][1, 2] == pytest.approx(noise, rel=1e-4)
    lam = np.sqrt(2 * 200 + np.sqrt(4 * 200)) * noise
    r = ob.fista(op, d, ob.complex_to_real(dirty), lam, noise, None, step=info0["step"])
    assert rel(F[1, :, 1, 2], ob.real_to_complex(r["x"])) < TOL_MODEL
    del order


def test_config1_reference_default_schedule_known_answers():
    """BASELINE configs[0] with the reference's DEFAULT schedule (README.md:157-210): lambda from the sigma heuristic,
    noise = theo_noise, 1458 iterations.  The numbers below were produced by the reference's own classes
    ("Reference-derived known answers"; 890 s per line of sight there)."""
    nu = np.linspace(1.008e9, 2.031e9, 1000)
    src = cs.FaradayThinSource(nu=nu, s_nu=0.0035, phi_gal=-200, spectral_idx=1.0)
    src.simulate()
    src.apply_noise(0.2 * 0.0035, random_state=np.random.RandomState(0))
    assert src.theo_noise == pytest.approx(2.213594283189171e-05, rel=1e-12)
    assert src.k == pytest.approx(2123786584.5300963, rel=1e-12)
    par = cs.Parameter()
    par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
    dft = cs.DFT1D(dataset=src, parameter=par)
    nufft = cs.NUFFT1D(dataset=src, parameter=par, solve=True)
    par.data = dft.backward(src.data)
    par.complex_data_to_real()
    lam = np.sqrt(src.m + 2 * np.sqrt(src.m)) * np.sqrt(2) * np.mean(src.sigma)
    assert lam == pytest.approx(0.03227972378805074, rel=1e-12)
    chi2, l1 = cs.Chi2(dft_obj=nufft), cs.L1(reg=lam)
    obj, X = cs.FISTA(guess_param=par, F_obj=cs.OFunction([chi2, l1]), fx=chi2, gx=cs.OFunction([l1]), noise=src.theo_noise,
                      verbose=False).run()
    X.real_data_to_complex()
    assert l1.reg == pytest.approx(5.519139152513642e-06, rel=1e-6)
    assert obj == pytest.approx(644.0708285925691, rel=2e-4)
    assert par.phi[np.argmax(np.abs(X.data))] == pytest.approx(-201.9582742926516)
    assert np.abs(X.data).max() == pytest.approx(0.019214807054993312, rel=1e-3)
    assert abs(np.count_nonzero(X.data) - 2719) <= 30  # the last iterations run at the edge of the unit step's stability
    assert np.sqrt(np.mean(np.abs(src.residual) ** 2)) == pytest.approx(0.00079447427061831, rel=1e-4)
