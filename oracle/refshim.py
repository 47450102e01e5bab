"""Import shim for the UNMODIFIED reference (test infrastructure, build container only).

The reference (`/root/reference/final_project`) cannot be imported as-is:
  * `final_project/arguments.py:13` lists a hard-coded `/Users/mac/...` directory at import,
  * `final_project/utils.py:7` imports matplotlib (absent) and `utils.py:37-38` opens a
    hard-coded calib.txt,
  * `final_project/algorithms/ransac.py:1` imports gtsam (absent) and `ransac.py:11` calls
    `read_cameras()` at import.

This shim stubs exactly those side effects (SURVEY.md section 8c recipe) and then imports the
reference modules from where they lie.  Nothing is copied out of /root/reference; the
reference source stays read-only.  `/root/reference` does not exist on the GPU box, so this
module is only ever used by `oracle/make_golden.py` and by CPU tests that skip when the
reference tree is absent.

Only `tests/`, `oracle/make_golden.py` may import this file.
"""
from __future__ import annotations

import builtins
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

# KITTI sequence-00 calibration rows (corroborated by
# VAN_ex/code/ex5.py:23).  The dataset itself is not shipped with the reference.
KITTI00_CALIB = (
    "P0: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 0.000000000000e+00 "
    "0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 "
    "0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00\n"
    "P1: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.861448000000e+02 "
    "0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 "
    "0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00\n"
)

_MAC_PREFIX = "/Users/mac/67604-SLAM-video-navigation"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "final_project"))


class _Pose3:
    """Minimal stand-in for gtsam.Pose3 (only used at ransac.py:199-200)."""

    def __init__(self, rot=None, t=None):
        self.R = np.eye(3) if rot is None else np.asarray(rot.R, dtype=np.float64)
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3)

    def inverse(self):
        out = _Pose3()
        out.R = self.R.T
        out.t = -self.R.T @ self.t
        return out

    def rotation(self):
        return _Rot3(self.R)

    def translation(self):
        return self.t

    def matrix(self):
        m = np.eye(4)
        m[:3, :3] = self.R
        m[:3, 3] = self.t
        return m


class _Rot3:
    def __init__(self, R=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)

    def matrix(self):
        return self.R


def _point3(v):
    return np.asarray(v, dtype=np.float64).reshape(3)


_loaded = None


def load(data_dir: str | None = None, n_frames: int = 4):
    """Import the reference hot-path modules through the shim.

    Returns a namespace with `.matching`, `.triangulation`, `.ransac`, `.utils`,
    `.tracking_database`, `.database` (the unmodified reference modules), with the module
    globals FEATURE / MATCHER_LEFT_RIGHT / MATCHER rebound to the AKAZE+Hamming configuration
    of `matching.py:19-24` (the checked-in default at `matching.py:72` is SIFT/L2).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)

    if data_dir is None:
        data_dir = tempfile.mkdtemp(prefix="slamfe_refshim_")
    seq = os.path.join(data_dir, "VAN_ex", "dataset", "sequences", "00")
    os.makedirs(os.path.join(seq, "image_0"), exist_ok=True)
    os.makedirs(os.path.join(seq, "image_1"), exist_ok=True)
    os.makedirs(os.path.join(data_dir, "VAN_ex", "dataset", "poses"), exist_ok=True)
    with open(os.path.join(seq, "calib.txt"), "w") as f:
        f.write(KITTI00_CALIB)
    with open(os.path.join(data_dir, "VAN_ex", "dataset", "poses", "00.txt"), "w") as f:
        for _ in range(n_frames):
            f.write("1 0 0 0 0 1 0 0 0 0 1 0\n")

    # (1) stub `final_project.arguments` (arguments.py:3-25) with paths inside data_dir
    args = types.ModuleType("final_project.arguments")
    args.MAC = True
    args.START = data_dir
    args.HEAD = data_dir + "/VAN_ex"
    args.DATA_PATH = args.HEAD + "/dataset/sequences/00/"
    args.LEN_DATA_SET = n_frames
    args.GROUND_TRUTH_PATH = args.HEAD + "/dataset/poses/00.txt"
    args.SIFT_DB_PATH = data_dir + "/SIFT_DB"
    args.AKAZE_DB_PATH = data_dir + "/AKAZE_DB"
    args.BUNDLES_PATH = data_dir + "/bundles_AKAZE"
    args.TRANSFORMATIONS_NPY = data_dir + "/pnp_global_transformations.npy"
    args.PNP_GLOBAL_T_PATH = args.TRANSFORMATIONS_NPY
    args.GRAPHS_DIR_PATH = data_dir + "/graphs"
    args.os = os
    sys.modules["final_project.arguments"] = args

    # (2) stub matplotlib, (3) stub gtsam
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "gtsam" not in sys.modules:
        g = types.ModuleType("gtsam")
        g.Pose3, g.Rot3, g.Point3 = _Pose3, _Rot3, _point3
        for name in ("utils", "symbol_shorthand"):
            sub = types.ModuleType("gtsam." + name)
            setattr(g, name, sub)
            sys.modules["gtsam." + name] = sub
        g.utils.plot = types.ModuleType("gtsam.utils.plot")
        sys.modules["gtsam.utils.plot"] = g.utils.plot

        class _Sym:
            def __call__(self, i):
                return i

        for s in "XLCPQKVB":
            setattr(g.symbol_shorthand, s, _Sym())
        sys.modules["gtsam"] = g

    # (4) redirect the hard-coded /Users/mac/... opens (utils.py:37-38) into data_dir
    real_open = builtins.open

    def redirected_open(path, *a, **k):
        if isinstance(path, str) and path.startswith(_MAC_PREFIX):
            path = data_dir + path[len(_MAC_PREFIX):]
        return real_open(path, *a, **k)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    builtins.open = redirected_open
    try:
        import final_project.algorithms.matching as matching
        import final_project.algorithms.triangulation as triangulation
        import final_project.utils as utils
        import final_project.algorithms.ransac as ransac
        import final_project.backend.database.tracking_database as tracking_database
        import final_project.backend.database.database as database
        import final_project.Inputs as inputs
    finally:
        builtins.open = real_open

    # AKAZE + Hamming configuration (matching.py:19-24), rebound at every by-value importer
    feat, lr, m = matching.get_akaze_matcher_lr_matcher()
    matching.FEATURE, matching.MATCHER_LEFT_RIGHT, matching.MATCHER = feat, lr, m
    database.MATCHER = m

    ns = types.SimpleNamespace(
        matching=matching, triangulation=triangulation, utils=utils, ransac=ransac,
        tracking_database=tracking_database, database=database, inputs=inputs,
        data_dir=data_dir, arguments=args)
    _loaded = ns
    return ns


def load_loop_closure():
    """Additionally import the reference's `backend/loop/loop_closure.py` (for
    `check_candidate_match`, :405-436).  Its import chain pulls in the GTSAM back-end modules
    (bundle.py, pose_graph.py, gtsam_utils.py), whose module-level code only needs gtsam NAMES to
    exist: the gtsam stub gets a permissive module `__getattr__` (any other attribute is a MagicMock).
    Nothing GTSAM-dependent is ever called through this shim."""
    from unittest.mock import MagicMock
    ns = load()
    if getattr(ns, "loop_closure", None) is not None:
        return ns
    for name in ("gtsam", "gtsam.utils", "gtsam.utils.plot", "gtsam.symbol_shorthand"):
        mod = sys.modules[name]
        if "__getattr__" not in mod.__dict__:
            mod.__dict__["__getattr__"] = lambda attr, _m=mod: MagicMock(name=f"{_m.__name__}.{attr}")
    real_open = builtins.open

    def redirected_open(path, *a, **k):
        if isinstance(path, str) and path.startswith(_MAC_PREFIX):
            path = ns.data_dir + path[len(_MAC_PREFIX):]
        return real_open(path, *a, **k)

    builtins.open = redirected_open
    try:
        import final_project.backend.loop.loop_closure as loop_closure
    finally:
        builtins.open = real_open
    loop_closure.MATCHER = ns.matching.MATCHER   # imported by value (loop_closure.py:12)
    ns.loop_closure = loop_closure
    return ns
