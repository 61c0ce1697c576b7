"""ctypes binding of ``libcsromer_b200.so`` (the C ABI declared in ``include/csromer_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``).
There is no CPU fallback: if the shared object is missing, or a compute entry point is called without an
sm_100 device, an exception is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcsromer_b200.so")

# every symbol include/csromer_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "csr_last_error", "csr_version", "csr_plan_create", "csr_plan_destroy", "csr_plan_dims",
    "csr_transform_ws_bytes", "csr_forward", "csr_adjoint", "csr_derotate", "csr_wavelet_create",
    "csr_wavelet_destroy", "csr_wavelet_ncoef", "csr_wavelet_levels", "csr_wavelet_level_sizes",
    "csr_wavelet_ws_bytes", "csr_wavelet_decompose", "csr_wavelet_reconstruct", "csr_fista_ws_bytes",
    "csr_fista", "csr_launch_count", "csr_profile", "csr_profile_collect", "csr_edge_noise", "csr_convolve_phi",
    "csr_peak", "csr_set_option", "csr_nudft_forward", "csr_nudft_adjoint", "csr_fista_nu_ws_bytes", "csr_fista_nu",
    "csr_pack_slab", "csr_unpack_slab", "csr_channel_stats", "csr_pack_planar", "csr_unpack_planar", "csr_sparse_count",
    "csr_sparse_fill", "csr_flag_mean", "csr_flag_hampel", "csr_galactic_rm_image", "csr_healpix_lookup",
    "csr_debug_timeline",
]


class CsrError(RuntimeError):
    pass


class FistaArgs(ctypes.Structure):
    """mirror of ``csr_fista_args``"""
    _fields_ = [
        ("P", c_int), ("iters", c_int), ("step", c_float), ("norm_factor", c_float), ("w_per_pixel", c_int),
        ("data", c_void_p), ("w", c_void_p), ("s", c_void_p), ("k", c_void_p), ("x", c_void_p), ("ldx", c_int),
        ("lambda0", c_void_p), ("noise", c_void_p), ("iter_cap", c_void_p), ("f_dirty", c_void_p), ("model", c_void_p),
        ("residual", c_void_p), ("cost", c_void_p), ("delta", c_void_p), ("use_dirty_as_guess", c_int),
        ("mma_blocks", c_void_p), ("s_per_pixel", c_int), ("f_dirty_in", c_void_p),
    ]


class Wcs(ctypes.Structure):
    """mirror of ``csr_wcs``"""
    _fields_ = [("crval1", c_double), ("crval2", c_double), ("crpix1", c_double), ("crpix2", c_double), ("cd11", c_double),
                ("cd12", c_double), ("cd21", c_double), ("cd22", c_double), ("lonpole", c_double), ("proj", c_int)]


class NuSampling(ctypes.Structure):
    """mirror of ``csr_nu_sampling``"""
    _fields_ = [("lambda2", c_void_p), ("ldl", c_int), ("m_count", c_void_p), ("m", c_int), ("n", c_int),
                ("phi0", c_double), ("dphi", c_double)]


_lib = None


def load():
    """Load the shared object (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CsrError(
            "csromer_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.csr_last_error.restype = c_char_p
    lib.csr_version.restype = c_int
    lib.csr_plan_create.argtypes = [POINTER(c_void_p), POINTER(c_double), c_int, c_double, POINTER(c_double), c_int,
                                    c_int]
    lib.csr_plan_destroy.argtypes = [c_void_p]
    lib.csr_plan_destroy.restype = None
    lib.csr_plan_dims.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                  POINTER(c_size_t)]
    lib.csr_transform_ws_bytes.argtypes = [c_void_p, c_int]
    lib.csr_transform_ws_bytes.restype = c_size_t
    for fn in (lib.csr_forward, lib.csr_adjoint):
        fn.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    lib.csr_derotate.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.csr_wavelet_create.argtypes = [POINTER(c_void_p), c_int, POINTER(c_double), POINTER(c_double),
                                       POINTER(c_double), POINTER(c_double), c_int, c_int, c_int, c_int, c_int, c_int]
    lib.csr_wavelet_destroy.argtypes = [c_void_p]
    lib.csr_wavelet_destroy.restype = None
    lib.csr_wavelet_ncoef.argtypes = [c_void_p]
    lib.csr_wavelet_levels.argtypes = [c_void_p]
    lib.csr_wavelet_level_sizes.argtypes = [c_void_p, POINTER(c_int), c_int]
    lib.csr_wavelet_ws_bytes.argtypes = [c_void_p, c_int]
    lib.csr_wavelet_ws_bytes.restype = c_size_t
    for fn in (lib.csr_wavelet_decompose, lib.csr_wavelet_reconstruct):
        fn.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]
    lib.csr_fista_ws_bytes.argtypes = [c_void_p, c_void_p, c_int]
    lib.csr_fista_ws_bytes.restype = c_size_t
    lib.csr_fista.argtypes = [c_void_p, c_void_p, POINTER(FistaArgs), c_void_p, c_size_t, c_void_p]
    lib.csr_edge_noise.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]
    lib.csr_convolve_phi.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_float), c_int, c_void_p, c_void_p, c_void_p]
    lib.csr_peak.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.csr_set_option.argtypes = [c_char_p, c_int]
    lib.csr_pack_slab.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.csr_unpack_slab.argtypes = [c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_void_p, c_void_p, c_void_p]
    lib.csr_channel_stats.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.csr_pack_planar.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
    lib.csr_unpack_planar.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_longlong, c_int, c_void_p]
    lib.csr_sparse_count.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.csr_sparse_fill.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]
    lib.csr_flag_mean.argtypes = [c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p]
    lib.csr_flag_hampel.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    lib.csr_galactic_rm_image.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(Wcs), c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p]
    lib.csr_healpix_lookup.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p]
    for fn in (lib.csr_nudft_forward, lib.csr_nudft_adjoint):
        fn.argtypes = [POINTER(NuSampling), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.csr_fista_nu_ws_bytes.argtypes = [POINTER(NuSampling), c_void_p, c_int]
    lib.csr_fista_nu_ws_bytes.restype = c_size_t
    lib.csr_fista_nu.argtypes = [POINTER(NuSampling), c_void_p, POINTER(FistaArgs), c_void_p, c_size_t, c_void_p]
    lib.csr_profile.argtypes = [c_int]
    lib.csr_profile_collect.argtypes = [POINTER(c_double), POINTER(c_longlong)]
    lib.csr_debug_timeline.argtypes = [POINTER(c_longlong), c_int]
    lib.csr_launch_count.argtypes = [c_int]
    lib.csr_launch_count.restype = c_longlong
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().csr_last_error().decode("utf-8", "replace")
        raise CsrError("%s failed (status %d): %s" % (what, rc, msg))


def set_option(name, value):
    """toggle one of the exact shortcuts (``zero_skip``, ``k_skip``, ``screen``, ``screen8``, ``pair``, ``sym``,
    ``resid_fast``, ``screen16_pair``, ``fused_wavelet``, ``wavelet_vec4``, ``wavelet_screen``, ``screen_incr``, ``rescale``,
    ``pair3``, ``tma_epi``, ``fwd128``, ``graph``: results do not change; ``nufft`` switches the per-line-of-sight transforms
    between the NUFFT and the direct sums, ``fwd_group`` sets how sparse forward tiles group their partial sums: last-bit changes)"""
    check(load().csr_set_option(name.encode(), int(value)), "csr_set_option")


def dptr(arr):
    return arr.ctypes.data_as(POINTER(c_double))
