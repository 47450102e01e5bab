"""slamfe — B200-native (sm_100a) stereo-SLAM front-end.

Drop-in for the data-parallel hot path of michaelpiro/67604-SLAM---video-navigation
(final_project/algorithms/{matching,triangulation,ransac}.py and the math bits of utils.py):
brute-force Hamming kNN matching of AKAZE descriptors with crossCheck / ratio test / rectified
stereo row filter, batched stereo DLT triangulation and RANSAC-PnP hypothesis scoring.

The directory name (`67604-slam---video-navigation_b200`) is not a Python identifier; import the
package as `slamfe` through the loader `slamfe.py` at the repository root.

Layout
    csrc/            hand-written CUDA kernels + the C-ABI (include/slamfe.h)
    _cabi.py         ctypes binding, ops.py  tensor-level operators
    matching.py / triangulation.py / ransac.py / utils.py   mirrors of the reference modules
    frontend.py      device-resident batched sequence pipeline (frame pairs per launch)
    loop.py          batched loop-closure candidate verification (match + RANSAC-PnP per candidate)
    database.py      batched drop-in for the reference's create_db loop (feeds TrackingDB.add_frame)
    dist.py          one-process-per-GPU sharding + NCCL all-gather of result tables
    patch.py         rebinding of the reference's module attributes (the "plugin" hook)
    synth.py         seeded synthetic KITTI-shaped inputs
"""
from . import _cabi  # noqa: F401  (ctypes prototypes; loading the .so is lazy)
from ._cabi import EXPORTED_SYMBOLS, KEY_IDX_BITS, KEY_IDX_MASK, KEY_NONE, SlamfeError, load_library

__version__ = "0.1.0"

__all__ = ["EXPORTED_SYMBOLS", "KEY_IDX_BITS", "KEY_IDX_MASK", "KEY_NONE", "SlamfeError", "load_library",
           "matching", "triangulation", "ransac", "utils", "ops", "frontend", "loop", "database", "dist", "patch", "synth"]


def __getattr__(name):
    if name in ("matching", "triangulation", "ransac", "utils", "ops", "frontend", "loop", "database", "dist", "patch",
                "synth", "build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
