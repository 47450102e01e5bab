"""cv2 container objects for the reference-typed drop-in, built in C (chost/objects.c).

`Matcher.match` / `.knnMatch` must hand back tuples of real cv2.DMatch (matching.py:44, database.py:54-55);
cv2's 4-argument constructor costs ~1 us per object, 3 ms for one frame's matches — more than the GPU call.
The helper calls the no-argument constructor and writes the four fields in place.  That relies on the
object layout PyObject_HEAD + cv::DMatch, so it is PROVEN here before use: objects made by the helper must
read back identical to objects made by cv2's own constructor, field by field, and cv2's own objects must
read back through the helper.  If the proof fails (another cv2 build), or the helper is not built, the
callers keep their pure-Python construction — this is object plumbing on the host, not a compute path.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

import numpy as np

_state = {"checked": False, "mod": None, "cls": None, "offset": 0}


def _load():
    if _state["checked"]:
        return _state["mod"]
    _state["checked"] = True
    try:
        import cv2
        from . import build as _build
        path = _build.build_objects()
        loader = importlib.machinery.ExtensionFileLoader("_slamfe_objects", path)
        spec = importlib.util.spec_from_loader("_slamfe_objects", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        cls = cv2.DMatch
        offset = cls.__basicsize__ - 16
        if offset < 16 or getattr(cls, "__itemsize__", 0) != 0:
            return None
        # proof 1: helper-made objects read back through cv2's getters
        q = np.array([1234567, 0, 2147483647], np.int32)
        t = np.array([7654321, 5, 0], np.int32)
        d = np.array([3.25, 0.0, 488.0], np.float32)
        made = mod.dmatch_tuple(cls, offset, q, t, d)
        for i, m in enumerate(made):
            if type(m) is not cls or (m.queryIdx, m.trainIdx, m.imgIdx, m.distance) != (int(q[i]), int(t[i]), 0, float(d[i])):
                return None
        # proof 2: cv2-made objects read back through the helper
        own = [cls(int(a), int(b), 0, float(c)) for a, b, c in zip(q, t, d)]
        qb, tb = mod.dmatch_indices(own, cls, offset)
        if not (np.array_equal(np.frombuffer(qb, np.int32), q) and np.array_equal(np.frombuffer(tb, np.int32), t)):
            return None
        _state.update(mod=mod, cls=cls, offset=offset)
    except Exception:
        _state["mod"] = None
    return _state["mod"]


def available() -> bool:
    return _load() is not None


def dmatch_tuple(qidx, tidx, dist):
    """tuple of cv2.DMatch(queryIdx, trainIdx, 0, distance), or None when the helper is not usable."""
    mod = _load()
    if mod is None:
        return None
    return mod.dmatch_tuple(_state["cls"], _state["offset"], np.ascontiguousarray(qidx, np.int32),
                            np.ascontiguousarray(tidx, np.int32), np.ascontiguousarray(dist, np.float32))


def knn_tuples(first, second, has_second):
    mod = _load()
    if mod is None:
        return None
    return mod.knn_tuples(first, second, np.ascontiguousarray(has_second, np.uint8))


def dmatch_indices(matches):
    """(queryIdx, trainIdx) int32 arrays of a sequence of cv2.DMatch, or None (helper unusable / other objects)."""
    mod = _load()
    if mod is None:
        return None
    try:
        qb, tb = mod.dmatch_indices(matches, _state["cls"], _state["offset"])
    except TypeError:
        return None
    return np.frombuffer(qb, np.int32), np.frombuffer(tb, np.int32)
