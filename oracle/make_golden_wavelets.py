"""Pin the wavelet restatement to PyWavelets itself -- when PyWavelets is importable.

    python oracle/make_golden_wavelets.py        -> tests/golden/wavelets_pywt.npz

PyWavelets 1.4.1 (the reference's requirements.txt:12) is a third-party dependency that is absent from /root/reference and
from this image (no network), which is why the wavelet parity of oracle/restatement.py is "unpinned" (DESIGN.md section 4).
This script is the pin: wherever `import pywt` works it runs the very calls the reference makes
(dictionaries/undecimated.py:75-82,145-159: pywt.swt(..., trim_approx=True, norm=...) / pywt.iswt;
dictionaries/discrete.py:37-40,80-89: pywt.wavedec / pywt.waverec with mode="periodization", pywt.ravel_coeffs) on seeded
inputs -- odd DWT lengths, norm=True, the reference's IUWT bank, coif2 at the C1 length -- and freezes inputs and outputs.
tests/test_oracle_wavelets.py::test_against_pywt_fixtures compares the restatement with the file when it exists and is
skipped otherwise.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wavelets_pywt.npz")
CASES = [  # (tag, kind, wavelet, N, level, norm)
    ("swt_coif2_15936", "swt", "coif2", 15936, None, False), ("swt_coif2_norm", "swt", "coif2", 3136, None, True),
    ("swt_sym2", "swt", "sym2", 3136, 4, False), ("swt_db4", "swt", "db4", 4096, None, False),
    ("swt_iuwt", "swt", "IUWT", 3136, None, False), ("dwt_db2_odd", "dwt", "db2", 3998, None, False),
    ("dwt_coif2", "dwt", "coif2", 15936, None, False), ("dwt_haar", "dwt", "db1", 1000, 5, False),
]


def main():
    try:
        import pywt
    except ImportError:
        print("PyWavelets is not importable here: no fixtures written (the wavelet pin stays 'unpinned')")
        return 1
    out = {"pywt_version": np.array(pywt.__version__)}
    rng = np.random.RandomState(2024)
    for tag, kind, name, N, level, norm in CASES:
        x = rng.normal(size=N)
        if name == "IUWT":  # the reference's own bank (dictionaries/undecimated.py:20-25)
            h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
            g = [-1 / 16, -4 / 16, 10 / 16, -4 / 16, -1 / 16]
            d = [0, 0, 1, 0, 0]
            wav = pywt.Wavelet("IUWT", filter_bank=[g, h, d, d])
        else:
            wav = pywt.Wavelet(name)
        if kind == "swt":
            lev = pywt.swt_max_level(N) if level is None else level
            coeffs = pywt.swt(x, wav, level=lev, trim_approx=True, norm=norm)
            rec = pywt.iswt(coeffs, wav, norm=norm)
        else:
            lev = pywt.dwt_max_level(N, wav.dec_len) if level is None else level
            coeffs = pywt.wavedec(x, wav, mode="periodization", level=lev)
            rec = pywt.waverec(coeffs, wav, mode="periodization")
        out[tag + "_x"] = x
        out[tag + "_flat"] = np.concatenate([np.asarray(c) for c in coeffs])
        out[tag + "_sizes"] = np.array([len(c) for c in coeffs])
        out[tag + "_rec"] = np.asarray(rec)
        out[tag + "_meta"] = np.array([kind, name, str(N), str(lev), str(int(norm))])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
