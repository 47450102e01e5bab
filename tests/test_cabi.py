"""The C-ABI library: loads, exports every symbol include/slamfe.h declares, rejects bad
arguments without touching the GPU.  CPU only (no compute calls)."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "slamfe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slamfe_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(slamfe):
    hdr = _header_symbols()
    assert hdr, "no declarations found in include/slamfe.h"
    assert sorted(slamfe.EXPORTED_SYMBOLS) == hdr


def test_library_exports_every_symbol(slamfe):
    lib = slamfe.load_library()
    for name in _header_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "67604-slam---video-navigation_b200",
                                                                     "libslamfe.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (slamfe_[a-z0-9_]+)", out))
    assert set(_header_symbols()) <= exported
    hdr_version = int(re.search(r"#define SLAMFE_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "slamfe.h")).read()).group(1))
    assert lib.slamfe_version() == hdr_version == slamfe._cabi.ABI_VERSION
    assert lib.slamfe_error_string(0) == b"ok"
    assert b"invalid" in lib.slamfe_error_string(-1)


def test_stale_library_is_detected(slamfe):
    """The loader compares a digest of csrc/ + include/slamfe.h with the one recorded at build time
    (a library older than its sources must not be called with a changed argument list)."""
    from slamfe import build
    assert os.path.exists(build.HASH_PATH) and not build.needs_build()
    assert open(build.HASH_PATH).read().strip() == build.source_hash()


def test_sass_is_sm100a_with_tma(slamfe):
    """The matcher must carry a bulk-copy (TMA) instruction and POPC in its sm_100a SASS; the default
    matcher must carry the tcgen05 MMA and TMEM load instructions."""
    so = os.path.join(ROOT, "67604-slam---video-navigation_b200", "libslamfe.so")
    res = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True)
    if res.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in res.stdout
    assert "UBLKCP" in res.stdout and "POPC" in res.stdout and "SYNCS" in res.stdout
    assert "UTCIMMA" in res.stdout and "LDTM" in res.stdout and "VIMNMX3.U16x2" in res.stdout


def test_argument_errors_do_not_need_a_gpu(slamfe):
    lib = slamfe.load_library()
    EINVAL, ERANGE = -1, -2
    buf = ctypes.create_string_buffer(1024)
    p = ctypes.addressof(buf)
    # negative sizes / null outputs / bad strides are rejected before any CUDA call
    assert lib.slamfe_hamming_top2(p, -1, 61, p, 4, 61, 61, 0, p, None, 0, None) == EINVAL
    assert lib.slamfe_hamming_top2(p, 4, 61, p, 4, 61, 61, 0, None, None, 0, None) == EINVAL
    assert lib.slamfe_hamming_top2(p, 4, 61, p, 4, 61, 65, 0, p, None, 0, None) == EINVAL
    assert lib.slamfe_hamming_top2(p, 4, 32, p, 4, 61, 61, 0, p, None, 0, None) == EINVAL
    assert lib.slamfe_hamming_top2(p, 4, 61, p, 1 << 23, 61, 61, 0, p, None, 0, None) == ERANGE
    assert lib.slamfe_hamming_top2(p, 0, 61, p, 4, 61, 61, 0, p, None, 0, None) == 0  # empty query: no-op
    assert lib.slamfe_unpack_keys(None, 5, p, p, None) == EINVAL
    assert lib.slamfe_unpack_keys(p, 0, p, p, None) == 0
    assert lib.slamfe_merge_top2(p, 2, -3, p, None) == EINVAL
    assert lib.slamfe_stereo_filter(p, p, p, p, -1, p, None) == EINVAL
    assert lib.slamfe_triangulate_dlt_f64(p, p, -1, p, p, p, None) == EINVAL
    P = (ctypes.c_double * 12)(*range(12))
    Q = (ctypes.c_double * 12)(*[v + (1 if i >= 4 else 0) for i, v in enumerate(range(12))])
    # rows 1-2 of P and Q differ: the links entry point refuses (general DLT must be used)
    assert lib.slamfe_triangulate_links_f64(p, 4, P, Q, p, None) == EINVAL
    assert lib.slamfe_ransac_score(None, None, 4, p, p, p, None, None, 5, 1, 5, None, None, None, p, p, p, p, None) == EINVAL
    assert lib.slamfe_peak_kernel(9, 0, 1, 1, p, None, 0, None) == EINVAL
