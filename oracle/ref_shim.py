"""Oracle A -- import the UNMODIFIED reference (``/root/reference/src/csromer``) through stub modules.

TEST INFRASTRUCTURE ONLY.  Nothing under ``csromer_b200/`` may import this file.  It exists so that
``oracle/make_golden.py`` can run the reference's own classes in the build container and freeze their
outputs under ``tests/golden/``; ``/root/reference`` does not exist on the GPU box; there the shim
imports the byte-for-byte copy ``oracle/build_ref.py`` leaves under ``oracle/_ref/`` (git-ignored, shipped with the
built artefacts), which is what bench.py's ``--impl reference`` arm and ``cpu_baseline`` time.

Why stubs are needed: ``astropy``, ``pywt``, ``pynufft`` and ``prox_tv`` are not
installed and there is no network.  The exact-DFT path (``Dataset``, ``Parameter``, ``NDFT1D``, ``Chi2``,
``L1``, ``OFunction``, ``FISTA``, ``simulation``) only touches them at import time, so empty stand-ins are
enough.  ``transformers/dfts/ndft.py:24`` reads ``self.k`` which does not exist; ``NDFT1DFixed`` supplies it
from ``dataset.k`` exactly as ``transformers/dfts/nufft.py:55`` does -- no reference file is modified.
"""
import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("CSROMER_REFERENCE_SRC", "/root/reference/src")
# the byte-for-byte run-time copy oracle/build_ref.py makes (git-ignored; travels to the GPU box)
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def source_dir():
    """directory that holds the reference's ``csromer`` package: the reference tree itself, else its copy"""
    for d in (REFERENCE_SRC, REF_COPY):
        if os.path.isdir(os.path.join(d, "csromer")):
            return d
    return None


def available():
    return source_dir() is not None


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    """Placeholder for classes the exact-DFT path never instantiates."""

    def __init__(self, *a, **k):
        raise RuntimeError("stubbed third-party class was instantiated; this path is not covered by Oracle A")


_loaded = None


def load():
    """Return a namespace with the reference's hot-path classes (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference sources not found at %s" % REFERENCE_SRC)

    if "astropy" not in sys.modules:
        astropy = _stub("astropy")
        astropy.units = _stub("astropy.units", rad=1.0)
        astropy.convolution = _stub("astropy.convolution", Gaussian1DKernel=_Anything)
        astropy.stats = _stub("astropy.stats", sigma_clipped_stats=_Anything)
    if "pywt" not in sys.modules:
        _stub("pywt", Wavelet=_Anything)
    if "pynufft" not in sys.modules:
        _stub("pynufft", NUFFT=_Anything)
    if "prox_tv" not in sys.modules:
        _stub("prox_tv")

    # bare package object: skips src/csromer/__init__.py (setuptools_scm version lookup)
    pkg = types.ModuleType("csromer")
    pkg.__path__ = [os.path.join(source_dir(), "csromer")]
    sys.modules["csromer"] = pkg
    # tv.py / tsv.py use an ndarray dataclass default, which python >= 3.11 rejects (tv.py:11, tsv.py:11)
    _stub("csromer.objectivefunction.priors.tv", TV=_Anything)
    _stub("csromer.objectivefunction.priors.tsv", TSV=_Anything)

    ns = types.SimpleNamespace()
    ns.Dataset = importlib.import_module("csromer.base.dataset").Dataset
    ns.Parameter = importlib.import_module("csromer.reconstruction.parameter").Parameter
    ndft = importlib.import_module("csromer.transformers.dfts.ndft")
    ns.NDFT1D = ndft.NDFT1D

    class NDFT1DFixed(ndft.NDFT1D):
        # supplies the attribute ndft.py:24 expects; same value nufft.py:55 uses
        k = property(lambda self: self.dataset.k)

    ns.NDFT1DFixed = NDFT1DFixed
    of = importlib.import_module("csromer.objectivefunction")
    ns.Chi2, ns.L1, ns.OFunction = of.Chi2, of.L1, of.OFunction
    ns.FISTA = importlib.import_module("csromer.optimization.methods.fista").FISTA
    sim = importlib.import_module("csromer.simulation")
    ns.FaradayThinSource = sim.FaradayThinSource
    ns.FaradayThickSource = sim.FaradayThickSource
    util = importlib.import_module("csromer.utils.utilities")
    ns.complex_to_real, ns.real_to_complex, ns.next_power_2 = (util.complex_to_real, util.real_to_complex,
                                                               util.next_power_2)
    _loaded = ns
    return ns
