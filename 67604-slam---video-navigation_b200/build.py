"""Build libslamfe.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

The shared library sits next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libslamfe.so")
SOURCES = ["hamming.cu", "hamming_mma.cu", "stereo.cu", "triangulate.cu", "ransac.cu", "ransac_gen.cu", "tracking.cu", "trackdb.cu", "refit.cu", "gating.cu", "peaks.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libslamfe.so cannot be built")


HASH_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """Digest of everything the library is built from (csrc/, include/slamfe.h, the flags).  Stored next to
    the library at build time: file times do not survive a snapshot copy, contents do."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    for path in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(REPO_ROOT, "include", "slamfe.h")]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    if not os.path.exists(HASH_PATH):   # a library of unknown provenance (digest not shipped): trust it
        return False
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(REPO_ROOT, "include"), "-I", CSRC]
    cmd += os.environ.get("SLAMFE_NVCC_DEFINES", "").split()   # development only, e.g. -DSLAMFE_MMA_DEV
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB_PATH


OBJECTS_SRC = os.path.join(PKG_DIR, "chost", "objects.c")


def objects_path() -> str:
    import sysconfig
    return os.path.join(PKG_DIR, "_slamfe_objects" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_objects(force: bool = False) -> str:
    """The host helper that builds cv2.DMatch tuples (chost/objects.c, plain C against Python.h): gcc, in-tree."""
    import hashlib
    import sysconfig
    out = objects_path()
    with open(OBJECTS_SRC, "rb") as fh:
        digest = hashlib.sha256(fh.read()).hexdigest()
    stamp = out + ".srchash"
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return out
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("gcc not found; the cv2 object helper cannot be built")
    tmp = out + f".tmp{os.getpid()}"
    cmd = [cc, "-O2", "-fPIC", "-shared", "-I", sysconfig.get_paths()["include"], OBJECTS_SRC, "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, out)   # atomic: ranks of a multi-process launch may build at the same time
    with open(stamp, "w") as fh:
        fh.write(digest + "\n")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_objects(force="--force" in sys.argv))
