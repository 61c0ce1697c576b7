"""diagnostic: error growth at the C5 sampling (4000 ch, n = 31968) for different bases / iteration counts"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

import csromer_b200 as cs  # noqa: E402
from oracle import restatement as ob  # noqa: E402
from test_gpu_parity import oracle_fista, oracle_operator, rel, run_fista  # noqa: E402

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.RandomState(1236)
P = 6
nu = np.linspace(0.9e9, 1.67e9, nch)
src = cs.FaradayThinSource(nu=nu, s_nu=0.0035 * rng.uniform(0.5, 1.5, size=P), phi_gal=rng.uniform(-500, 500, size=P),
                           spectral_idx=1.0)
src.simulate()
src.apply_noise(0.2 * 0.0035, random_state=rng)
par = cs.Parameter()
par.calculate_cellsize(dataset=src, oversampling=8, verbose=False)
print("m", src.m, "n", len(par.phi), flush=True)
for label, wav, owav, Ks in (("delta", None, None, (2, 6, 20)),
                             ("swt coif2", cs.UndecimatedWavelet(wavelet_name="coif2"), ob.WaveletDict("coif2", kind="swt"), (1, 2, 6)),
                             ("swt coif2 norm", cs.UndecimatedWavelet(wavelet_name="coif2", norm=True), ob.WaveletDict("coif2", kind="swt", norm=True), (2, 6))):
    for K in Ks:
        cost, X, l1, lam0, noise = run_fista(src, par, K, wav=wav)
        ref = oracle_fista(src, par, K, lam0, noise, owav=owav)
        print("%-16s K=%2d err(x)=%.2e err(model)=%.2e err(cost)=%.2e  chi2-ish cost=%.4g  max|x|=%.3g" % (
            label, K, rel(X.data, ref["x"]), rel(src.model_data, ref["model"]), rel(cost, ref["cost"]), float(np.mean(ref["cost"])),
            np.abs(ref["x"]).max()), flush=True)
