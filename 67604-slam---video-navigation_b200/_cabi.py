"""ctypes binding of libslamfe.so (the C-ABI declared in include/slamfe.h).

No CPU fallback: if the library or a CUDA device is missing, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int64, c_uint64, c_void_p, POINTER

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SLAMFE_LIBRARY overrides the in-tree build (development: A/B timing of two builds on one box)
LIB_PATH = os.environ.get("SLAMFE_LIBRARY") or os.path.join(PKG_DIR, "libslamfe.so")

KEY_IDX_BITS = 22
KEY_IDX_MASK = (1 << KEY_IDX_BITS) - 1
KEY_NONE = 0xFFFFFFFF
MAX_DESC_BYTES = 64
MATCH_BEST_ONLY = 1
MATCH_COMPACT_KEYS = 2
MATCH_MMA = 4
ABI_VERSION = 205   # SLAMFE_ABI_VERSION of include/slamfe.h this binding was written against

# name -> (restype, argtypes); mirrors include/slamfe.h one to one
_SIGNATURES = {
    "slamfe_version": (c_int, []),
    "slamfe_error_string": (c_char_p, [c_int]),
    "slamfe_hamming_top2": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p]),
    "slamfe_hamming_top2_batched": (c_int, [c_void_p, c_int, c_void_p, c_void_p,
                                            c_void_p, c_int, c_void_p, c_void_p,
                                            c_int, c_int, c_int, c_int, c_int,
                                            c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "slamfe_hamming_top2_pairs": (c_int, [c_void_p, c_int, c_void_p, c_void_p,
                                          c_void_p, c_int, c_void_p, c_void_p,
                                          c_void_p, c_int, c_int, c_int, c_int,
                                          c_void_p, c_int64, c_int, c_void_p]),
    "slamfe_unpack_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "slamfe_merge_top2": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "slamfe_cross_check": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "slamfe_ratio_test": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "slamfe_stereo_filter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "slamfe_stereo_links_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                            c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "slamfe_triangulate_links_f64": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_triangulate_links_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_triangulate_dlt_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_ransac_score": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_track_gather": (c_int, [c_void_p] * 10 + [c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 9),
    "slamfe_pairs_gather": (c_int, [c_void_p] * 5 + [c_int, c_int] + [c_void_p] * 6),
    "slamfe_scatter_inliers": (c_int, [c_void_p] * 5 + [c_int, c_void_p, c_int64, c_void_p]),
    "slamfe_ransac_hypotheses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_track_ids": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64] + [c_void_p] * 7),
    "slamfe_pack_db": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_int] + [c_void_p] * 7),
    "slamfe_gate_candidates": (c_int, [c_void_p, c_int] + [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "slamfe_pnp_refit": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_void_p, c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slamfe_peak_kernel": (c_int, [c_int, c_int, c_int, c_int, c_void_p, POINTER(c_int), c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class SlamfeError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """dlopen libslamfe.so and set the prototypes.  Needs no GPU (symbols only)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("SLAMFE_LIBRARY"):
        # the in-tree build must match csrc/ and include/slamfe.h (a stale library would be called with the
        # wrong argument lists): rebuild when the recorded source digest differs.  Under a multi-process
        # launch a rebuild would race between ranks, so a stale library is an error there.
        from . import build as _build
        if _build.needs_build():
            if not build_if_missing and not os.path.exists(LIB_PATH):
                raise SlamfeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build`")
            if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.path.exists(LIB_PATH):
                raise SlamfeError(f"{LIB_PATH} is older than its sources: run `python __graft_entry__.py build` "
                                  f"before a multi-process launch")
            _build.build(force=True)
    elif not os.path.exists(LIB_PATH):
        raise SlamfeError(f"SLAMFE_LIBRARY={LIB_PATH} does not exist")
    lib = ctypes.CDLL(LIB_PATH)
    lib.slamfe_version.restype = c_int
    if lib.slamfe_version() != ABI_VERSION:
        raise SlamfeError(f"{LIB_PATH} has ABI version {lib.slamfe_version()}, this binding expects {ABI_VERSION}: "
                          f"run `python __graft_entry__.py build`")
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load_library().slamfe_error_string(rc).decode()
        raise SlamfeError(f"{what or 'libslamfe'} failed: {msg} (code {rc})")


def require_cuda():
    """The product path is CUDA-only: fail loudly instead of falling back to the CPU."""
    import torch
    if not torch.cuda.is_available():
        raise SlamfeError("slamfe needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, 0 for None.  An EMPTY view (zero rows of a larger
    table, e.g. the rows of a frame without keypoints) still gets the address it would start at —
    torch reports 0 for it, which the C-ABI would reject as a missing argument."""
    if t is None:
        return 0
    p = t.data_ptr()
    if p == 0 and t.numel() == 0:
        st = t.untyped_storage()
        if st.nbytes() > 0:
            p = st.data_ptr() + t.storage_offset() * t.element_size()
    return p


def host_doubles(a, n: int):
    """Small host-side double array (camera matrices) as a ctypes buffer."""
    import numpy as np
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if arr.size != n:
        raise ValueError(f"expected {n} doubles, got {arr.size}")
    buf = (ctypes.c_double * n)(*arr.tolist())
    return buf


def stream_handle(stream=None) -> int:
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return s.cuda_stream
