"""Triangulation kernels against the reference's np.linalg.svd DLT (golden vectors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-9   # fp64 kernels vs LAPACK: both are accurate to ~1e-13, bound is conservative
RTOL_F32 = 1e-5   # north_star tolerance for the fp32 path


def _rel(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


class _Link:
    def __init__(self, xl, xr, y):
        self.x_left, self.x_right, self.y = xl, xr, y


def test_links_f64_and_dropin(slamfe, golden):
    from slamfe import triangulation
    g = golden("triangulation")
    links = [_Link(*r) for r in g["links"]]
    xyz = triangulation.triangulate_links(links, g["P"], g["Q"])
    assert xyz.shape == g["xyz"].shape and xyz.dtype == np.float64
    assert _rel(xyz, g["xyz"]).max() < RTOL_F64

    class _DB:
        def all_last_frame_links(self):
            return links
    assert np.array_equal(triangulation.triangulate_last_frame(_DB(), g["P"], g["Q"]), xyz)
    assert np.array_equal(triangulation.triangulate_last_frame(None, g["P"], g["Q"], links=links[:7]), xyz[:7])
    assert triangulation.triangulate_links([], g["P"], g["Q"]).shape == (0, 3)


def test_links_f32(slamfe, golden):
    import torch
    from slamfe import ops
    g = golden("triangulation")
    xyz = ops.triangulate_links(torch.from_numpy(g["links"].astype(np.float32)).cuda(), g["P"], g["Q"]).cpu().numpy()
    assert xyz.dtype == np.float32
    assert _rel(xyz.astype(np.float64), g["xyz"]).max() < RTOL_F32


def test_general_dlt(slamfe, golden):
    from slamfe import triangulation
    g = golden("triangulation")
    xyz = triangulation.triangulate_points(g["P"], g["Q"], g["pxy"], g["qxy"])
    assert _rel(xyz, g["xyz_dlt"]).max() < RTOL_F64
    xyz2 = triangulation.triangulate_points(g["P"], g["Q2"], g["pxy2"], g["qxy2"])
    assert _rel(xyz2, g["xyz_gen"]).max() < RTOL_F64
    one = triangulation.linear_least_squares_triangulation(g["P"], g["Q"], tuple(g["pxy"][3]), tuple(g["qxy"][3]))
    assert one.shape == (3,) and np.allclose(one, g["xyz_dlt"][3], rtol=RTOL_F64)
    # links through a non-rectified pair take the general path inside triangulate_links
    links = np.stack([g["pxy2"][:, 0], g["qxy2"][:, 0], g["pxy2"][:, 1]], axis=1)
    got = triangulation.triangulate_links(links, g["P"], g["Q2"])
    ref = np.array([np.linalg.svd(_dlt(g["P"], g["Q2"], r))[2][-1] for r in links])
    assert _rel(got, ref[:, :3] / ref[:, 3:]).max() < 1e-7


def _dlt(P, Q, r):
    xl, xr, y = r
    return np.stack([P[2] * xl - P[0], P[2] * y - P[1], Q[2] * xr - Q[0], Q[2] * y - Q[1]])


def test_large_random_vs_oracle(slamfe, oracle):
    from slamfe import ransac, synth, triangulation
    rng = np.random.default_rng(41)
    links = synth.links(rng, 3000)
    got = triangulation.triangulate_links(links, ransac.P, ransac.Q)
    assert _rel(got, oracle.triangulate_links(links, ransac.P, ransac.Q)).max() < RTOL_F64
