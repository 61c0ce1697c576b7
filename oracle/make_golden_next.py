"""Generate tests/golden/next_rows.npz: outputs of the UNMODIFIED reference for the rows next to the hot path
(SURVEY.md section 8f): the flaggers (transformers/flaggers/*.py), ``calculate_noise`` (utils/utilities.py:68-110) and
``filter_cubes`` (io/io_functions.py:15-33).  TEST INFRASTRUCTURE ONLY; run once in the build container
(``python oracle/make_golden_next.py``) -- ``/root/reference`` does not travel to the GPU box.

``io_functions.py`` imports ``dask.array`` and ``astropy.io.fits`` at module level (absent here, not used by
``filter_cubes``): they are stubbed like the other third-party imports in ``oracle/ref_shim.py``.
"""
import importlib
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "next_rows.npz")


def main():
    R = ref_shim.load()
    for name in ("dask", "dask.array", "astropy.io", "astropy.io.fits"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["dask"].array = sys.modules["dask.array"]
    sys.modules["astropy.io"].fits = sys.modules["astropy.io.fits"]
    fl = importlib.import_module("csromer.transformers.flaggers")
    util = importlib.import_module("csromer.utils.utilities")
    iof = importlib.import_module("csromer.io.io_functions")

    rng = np.random.RandomState(42)
    m = 120
    nu = np.linspace(1.008e9, 2.031e9, m)
    sigma = 0.001 * (1.0 + 0.2 * rng.uniform(size=m))
    sigma[[7, 30, 31, 90]] *= 6.0           # RFI-like channels
    data = (rng.normal(size=m) + 1j * rng.normal(size=m)).astype(np.complex64)
    out = dict(nu=nu, sigma=sigma, data=data)

    def fresh():
        return R.Dataset(nu=nu, sigma=sigma.copy(), data=data.copy(), spectral_idx=0.7)

    for tag, nsig, delete in (("mean_ns0_w", 0.0, False), ("mean_ns5_w", 5.0, False), ("mean_ns5_del", 5.0, True)):
        ds = fresh()
        idxs, outl = fl.MeanFlagger(dataset=ds, nsigma=nsig, delete_channels=delete).run()
        out.update({tag + "_idxs": idxs, tag + "_outliers": outl, tag + "_w": ds.w, tag + "_lambda2": ds.lambda2,
                    tag + "_sigma": ds.sigma, tag + "_data": ds.data, tag + "_k": np.float64(ds.k)})
    hampel = importlib.import_module("csromer.transformers.flaggers.hampel_flagger").hampel
    keep, outl = hampel(sigma, 9, 3.0, False)
    imputed, keep2, outl2 = hampel(sigma, 9, 3.0, True)
    out.update(hampel_keep=keep, hampel_outliers=outl, hampel_imputed=imputed)
    ds = fresh()
    man = fl.ManualFlagger(dataset=ds, delete_channels=True)
    man.outlier_idxs = [3, 4, 50]
    idxs, _ = man.run()
    out.update(manual_idxs=idxs, manual_lambda2=ds.lambda2, manual_sigma=ds.sigma, manual_data=ds.data)
    out["mad"] = np.float64(importlib.import_module("csromer.transformers.flaggers.flagger").median_absolute_deviation(sigma))

    cube = rng.normal(size=(6, 20, 24)).astype(np.float32) * np.arange(1, 7)[:, None, None]
    cube[2, 3:5, 4:9] = np.nan
    cube[4] = 0.0
    out["cube"] = cube
    out["noise_full"] = util.calculate_noise(image=cube)
    out["noise_window"] = util.calculate_noise(image=cube, x0=2, xn=20, y0=1, yn=15)
    out["noise_image"] = np.float64(util.calculate_noise(image=cube[1]))
    header = {"CRVAL3": 1.0e9, "NAXIS3": 6, "CDELT3": 1.0e6}
    zero = np.zeros_like(cube)
    zero[4] = 0.0
    fI, fQ, fU, fnu = iof.filter_cubes(zero.copy(), cube, cube * 0.5, header)
    out.update(filter_Q=fQ, filter_nu=fnu)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.shape(v) for k, v in out.items() if k.startswith(("mean_ns5_del", "noise", "filter"))})


if __name__ == "__main__":
    main()
