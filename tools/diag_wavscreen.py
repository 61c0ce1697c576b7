"""diagnostic: wavelet-domain screening on / off, where do the results differ?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import csromer_b200 as cs
from csromer_b200 import _lib
import test_gpu_parity as T

nch, P, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
src, par = T.make_batch(nch, P, seed=78, mixed=True)
res = {}
for opt in ("on", "off"):
    _lib.set_option(os.environ.get("DIAG_OPT", "wavelet_screen"), 1 if opt == "on" else 0)
    for k in ([K] if len(sys.argv) < 5 else range(1, K + 1)):
        wav = cs.UndecimatedWavelet(wavelet_name="coif2")
        cost, X, l1, lam0, noise = T.run_fista(src, par, k, wav=wav, step=1.0 / 2.8)
        res[(opt, k)] = np.array(X.data, copy=True)
for k in ([K] if len(sys.argv) < 5 else range(1, K + 1)):
    a, b = res[("on", k)], res[("off", k)]
    N = 2 * len(par.phi)
    bad = np.argwhere(a != b)
    print("K", k, "shape", a.shape, "mismatches", len(bad), "max abs diff", np.abs(a - b).max(), "max", np.abs(b).max(),
          "nnz on/off", np.count_nonzero(a), np.count_nonzero(b))
    for (i, p) in bad[:12]:
        print("   coef", i, "band", i // N, "pos", i % N, "LOS", p, "on", a[i, p], "off", b[i, p])
    if len(bad):
        break
