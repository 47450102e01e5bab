"""The C helper that builds cv2.DMatch tuples for the reference-typed drop-in (slamfe/_objects.py,
chost/objects.c): same objects as cv2's own constructor, field by field (matching.py:44,
database.py:54-55 consume them), and a clean refusal for anything that is not a cv2.DMatch."""
import gc
import sys

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import slamfe  # noqa: E402
from slamfe import _objects, matching  # noqa: E402


def test_helper_is_built_and_proven():
    assert _objects.available(), "gcc + Python.h are part of the image: the helper must build and pass its layout proof"


def test_dmatch_tuple_equals_cv2_constructor():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 3000):
        q = rng.integers(0, 2**31 - 1, n).astype(np.int32)
        t = rng.integers(0, 2**31 - 1, n).astype(np.int32)
        d = rng.integers(0, 489, n).astype(np.int32)   # Hamming distances arrive as int32, leave as float
        got = matching._dmatches(q, t, d)
        assert isinstance(got, tuple) and len(got) == n
        for i in rng.permutation(n)[:200]:
            ref = cv2.DMatch(int(q[i]), int(t[i]), 0, float(d[i]))
            m = got[i]
            assert type(m) is cv2.DMatch
            assert (m.queryIdx, m.trainIdx, m.imgIdx, m.distance) == (ref.queryIdx, ref.trainIdx, ref.imgIdx, ref.distance)
            assert isinstance(m.distance, float)


def test_objects_survive_the_arrays_and_are_mutable():
    q = np.arange(5, dtype=np.int32)
    got = _objects.dmatch_tuple(q, q + 10, q.astype(np.float32) * 0.5)
    del q
    gc.collect()
    assert [m.trainIdx for m in got] == [10, 11, 12, 13, 14]
    got[2].trainIdx = 99            # plain cv2 objects: callers may edit them
    assert got[2].trainIdx == 99
    ref = (cv2.DMatch(0, 0, 0, 0.0), cv2.DMatch(1, 1, 0, 0.0), cv2.DMatch(2, 2, 0, 0.0))
    assert sys.getrefcount(got[2]) == sys.getrefcount(ref[2])   # one owner: the tuple


def test_knn_tuples_shape():
    first = _objects.dmatch_tuple(np.arange(4, dtype=np.int32), np.arange(4, dtype=np.int32), np.zeros(4, np.float32))
    has2 = np.array([True, False, True, True])
    second = _objects.dmatch_tuple(np.array([0, 2, 3], np.int32), np.array([7, 8, 9], np.int32), np.ones(3, np.float32))
    out = _objects.knn_tuples(first, second, has2)
    assert [len(x) for x in out] == [2, 1, 2, 2]
    assert out[0][0] is first[0] and out[2][1] is second[1] and out[3][1].trainIdx == 9
    with pytest.raises(ValueError):
        _objects.knn_tuples(first, second[:1], has2)


def test_dmatch_indices_roundtrip_and_refusal():
    rng = np.random.default_rng(1)
    q = rng.integers(0, 5000, 300).astype(np.int32)
    t = rng.integers(0, 5000, 300).astype(np.int32)
    own = [cv2.DMatch(int(a), int(b), 0, 1.0) for a, b in zip(q, t)]     # cv2-made objects, a list
    gq, gt = _objects.dmatch_indices(own)
    assert np.array_equal(gq, q) and np.array_equal(gt, t)

    class Fake:   # duck-typed matches (the tests of the drop-in use them): the helper refuses, the caller falls back
        queryIdx, trainIdx = 1, 2
    assert _objects.dmatch_indices([Fake(), Fake()]) is None
    assert _objects.dmatch_indices(own + [Fake()]) is None
