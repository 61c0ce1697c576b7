"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/csromer_b200.h declares
(no compute calls: this file runs without a GPU)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT

from csromer_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "csromer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(built_library):
    for name in declared_symbols():
        assert hasattr(built_library, name), name
    assert built_library.csr_version() >= 100


def test_fista_args_struct_layout(built_library):
    # pointer-sized fields after five 4-byte scalars (+4 padding); catches drift between header and ctypes mirror
    a = _lib.FistaArgs
    assert a.data.offset == 24 and a.x.offset == 56 and a.ldx.offset == 64
    assert a.iter_cap.offset == a.noise.offset + 8 and a.use_dirty_as_guess.offset == a.delta.offset + 8
    header = open(os.path.join(ROOT, "include", "csromer_b200.h")).read()
    body = header[header.index("typedef struct {"):header.index("} csr_fista_args;")]
    names = re.findall(r"(\w+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == [f[0] for f in a._fields_]


def test_error_paths_without_device(built_library):
    lib = built_library
    assert lib.csr_transform_ws_bytes(None, 10) == 0
    assert lib.csr_fista_ws_bytes(None, None, 10) == 0
    rc = lib.csr_plan_create(None, None, 0, 0.0, None, 0, 0)
    assert rc == -1 and b"null" in lib.csr_last_error()
    handle = ctypes.c_void_p()
    z = (ctypes.c_double * 4)(0.1, 0.2, 0.3, 0.4)
    # decimated wavelet object needs no device: sizes follow pywt's periodization rule
    rc = lib.csr_wavelet_create(ctypes.byref(handle), 1, z, z, z, z, 4, 3, 250, 0, 0, 0)
    assert rc == 0
    sizes = (ctypes.c_int * 8)()
    assert lib.csr_wavelet_level_sizes(handle, sizes, 8) == 4
    assert list(sizes)[:4] == [32, 32, 63, 125] and lib.csr_wavelet_ncoef(handle) == 252
    lib.csr_wavelet_destroy(handle)
    rc = lib.csr_wavelet_create(ctypes.byref(handle), 0, z, z, z, z, 3, 2, 64, 0, 0, 0)
    assert rc == -1


def test_documented_options_are_accepted(built_library):
    """every run-time switch the header documents is a name csr_set_option knows (and unknown names are refused)"""
    lib = built_library
    header = open(os.path.join(ROOT, "include", "csromer_b200.h")).read()
    doc = header[header.index("run-time switches of the exact shortcuts"):header.index("int csr_set_option")]
    names = re.findall(r'"(\w+)"', doc)
    assert {"zero_skip", "k_skip", "screen", "screen8", "pair", "sym", "resid_fast", "screen16_pair", "fused_wavelet",
            "wavelet_vec4", "wavelet_screen", "screen_incr", "rescale", "pair3", "tma_epi", "fwd128", "nufft", "graph", "fwd_group", "fwd16"} <= set(names)
    for name in names:
        assert lib.csr_set_option(name.encode(), 1) == 0, name
    assert lib.csr_set_option(b"no_such_option", 1) != 0 and b"unknown option" in lib.csr_last_error()


def test_sass_contains_cta_pair_and_packed_fma_instructions(built_library):
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "FFMA2" in sass, "packed fp32 FMAs (fused wavelet kernels) missing from SASS"
    assert "UTCHMMA.2CTA" in sass and "UTCQMMA.2CTA" in sass, "cta_group::2 MMAs (fp16 / fp8 pair kernels) missing from SASS"


def test_sass_contains_blackwell_tensor_and_tma_instructions(built_library):
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "UTMASTG" in sass, "TMA stores (staged residual epilogue of dft_gemm_pair_kernel) missing from SASS"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
