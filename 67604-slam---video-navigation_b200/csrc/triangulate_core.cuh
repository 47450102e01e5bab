// Arithmetic core of the link triangulation (used by triangulate.cu and tracking.cu; plain C++ under
// SLAMFE_HD so that oracle/triangulate_host_shim.cpp can compile it for the host and the CPU suite can
// check it against the reference's np.linalg.svd path on extreme inputs).
#pragma once
#include <math.h>

#include "hd.cuh"

namespace slamfe {

struct Cams {
    double P[12];
    double Q[12];
};

// Null vector of the rank-3 DLT system of a link; returns X = n[:3] / n[3].
// Rows: a = P[2]*xl - P[0] (triangulation.py:17), b = P[2]*y - P[1] (:18, == :20 for links),
// d = Q[2]*xr - Q[0] (:19).  n = generalized cross product of (a, b, d): the six 2x2 minors of
// (b, d) are shared by the four cofactors (36 fp64 operations + one reciprocal per link).
SLAMFE_HD void triangulate_link(const Cams &c, double xl, double xr, double y, double &X, double &Y,
                                                 double &Z)
{
    double a[4], b[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = fma(c.P[8 + k], xl, -c.P[k]);
        b[k] = fma(c.P[8 + k], y, -c.P[4 + k]);
        d[k] = fma(c.Q[8 + k], xr, -c.Q[k]);
    }
    const double m01 = b[0] * d[1] - b[1] * d[0], m02 = b[0] * d[2] - b[2] * d[0], m03 = b[0] * d[3] - b[3] * d[0];
    const double m12 = b[1] * d[2] - b[2] * d[1], m13 = b[1] * d[3] - b[3] * d[1], m23 = b[2] * d[3] - b[3] * d[2];
    const double n0 = a[1] * m23 - a[2] * m13 + a[3] * m12;
    const double n1 = -(a[0] * m23 - a[2] * m03 + a[3] * m02);
    const double n2 = a[0] * m13 - a[1] * m03 + a[3] * m01;
    double n3 = -(a[0] * m12 - a[1] * m02 + a[2] * m01);
    double s = 1.0;
    if (n3 == 0.0) {  // triangulation.py:22-23 guard on the unit-norm singular vector
        s = 1.0 / sqrt(n0 * n0 + n1 * n1 + n2 * n2);
        n3 = 1e-20;
    }
    const double inv = s / n3;
    X = n0 * inv;
    Y = n1 * inv;
    Z = n2 * inv;
}

}  // namespace slamfe
