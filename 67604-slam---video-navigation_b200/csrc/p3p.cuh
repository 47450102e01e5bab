// Minimal-sample pose solver for RANSAC-PnP hypothesis generation (fp64, one thread per sample).
//
// Replaces the host-side `cv2.solvePnP(points_3d[idx], pixels[idx], K, flags=SOLVEPNP_EPNP)` on 4
// sampled correspondences (final_project/algorithms/ransac.py:95-104, :157-171) + `rodriguez_to_mat`
// (final_project/utils.py:16-18).  OpenCV's EPnP on exactly 4 points is implementation-defined (the
// null space of its 12x12 moment matrix is 4-dimensional) and the reference samples with an unseeded
// RNG, so this is NOT a bit-for-bit replacement: it solves the exact minimal problem instead —
// P3P (Grunert: law of cosines -> quartic in v = s3/s1, coefficients derived with sympy) on the first
// three points, up to four poses, and the fourth point picks the pose with the smallest left-image
// reprojection error (the design of cv2.SOLVEPNP_P3P, against which oracle/p3p_oracle.py is pinned).
// Contract and tests: DESIGN.md section 2.6.
//
// The header is plain C++ under SLAMFE_HD so that oracle/p3p_host_shim.cpp can compile the very same
// code for the host in the CPU test suite; the product only ever runs it inside ransac_gen.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#include "hd.cuh"

namespace slamfe {
namespace p3p {

// Largest real root of t^3 + B t^2 + C t + D = 0 (Cardano / trigonometric form + Newton polish).
SLAMFE_HD double cubic_largest_root(double B, double C, double D)
{
    const double P = C - B * B / 3.0;
    const double Q = 2.0 * B * B * B / 27.0 - B * C / 3.0 + D;
    const double disc = 0.25 * Q * Q + P * P * P / 27.0;
    double z;
    if (disc > 0.0) {
        const double sq = sqrt(disc);
        z = cbrt(-0.5 * Q + sq) + cbrt(-0.5 * Q - sq);
    } else if (P < 0.0) {
        const double m = 2.0 * sqrt(-P / 3.0);
        double arg = 3.0 * Q / (P * m);
        arg = fmin(1.0, fmax(-1.0, arg));
        z = m * cos(acos(arg) / 3.0);
    } else {
        z = 0.0;
    }
    double t = z - B / 3.0;
    for (int it = 0; it < 3; ++it) {
        const double f = ((t + B) * t + C) * t + D;
        const double df = (3.0 * t + 2.0 * B) * t + C;
        if (df != 0.0) t -= f / df;
    }
    return t;
}

// Real roots of A4 x^4 + A3 x^3 + A2 x^2 + A1 x + A0 = 0 (Ferrari via the resolvent cubic, each root
// polished with Newton steps on the original quartic).  Returns the number of roots written.
SLAMFE_HD int quartic_real_roots_monic_side(double A4, double A3, double A2, double A1, double A0, double (&x)[4])
{
    if (A4 == 0.0 || !(fabs(A4) > 0.0)) return 0;
    const double b = A3 / A4, c = A2 / A4, d = A1 / A4, e = A0 / A4;
    const double b2 = b * b;
    const double p = c - 0.375 * b2;
    const double q = d - 0.5 * b * c + 0.125 * b2 * b;
    const double r = e - 0.25 * b * d + 0.0625 * b2 * c - 0.01171875 * b2 * b2;
    const double shift = -0.25 * b;
    int n = 0;
    const double scale = fabs(p) + sqrt(fabs(r)) + 1e-300;
    if (fabs(q) <= 1e-14 * scale * sqrt(scale)) {  // biquadratic
        const double disc = p * p - 4.0 * r;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double y2a = 0.5 * (-p + sq), y2b = 0.5 * (-p - sq);
            if (y2a >= 0.0) { const double y = sqrt(y2a); x[n++] = y + shift; x[n++] = -y + shift; }
            if (y2b >= 0.0) { const double y = sqrt(y2b); x[n++] = y + shift; x[n++] = -y + shift; }
        }
    } else {
        // resolvent: m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, m > 0
        const double m = cubic_largest_root(p, 0.25 * p * p - r, -0.125 * q * q);
        if (!(m > 0.0)) return 0;
        const double s = sqrt(2.0 * m);
        const double t1 = -2.0 * p - 2.0 * m - 2.0 * q / s;
        const double t2 = -2.0 * p - 2.0 * m + 2.0 * q / s;
        if (t1 >= 0.0) { const double w = sqrt(t1); x[n++] = 0.5 * (s + w) + shift; x[n++] = 0.5 * (s - w) + shift; }
        if (t2 >= 0.0) { const double w = sqrt(t2); x[n++] = 0.5 * (-s + w) + shift; x[n++] = 0.5 * (-s - w) + shift; }
    }
    int kept = 0;
    for (int k = 0; k < n; ++k) {
        double v = x[k];
        for (int it = 0; it < 3; ++it) {
            const double f = (((A4 * v + A3) * v + A2) * v + A1) * v + A0;
            const double df = ((4.0 * A4 * v + 3.0 * A3) * v + 2.0 * A2) * v + A1;
            if (df != 0.0) v -= f / df;
        }
        // keep only genuine roots: residual small against the magnitude of the terms
        const double v2 = v * v;
        const double f = (((A4 * v + A3) * v + A2) * v + A1) * v + A0;
        const double mag = fabs(A4) * v2 * v2 + fabs(A3 * v) * v2 + fabs(A2) * v2 + fabs(A1 * v) + fabs(A0);
        if (fabs(f) <= 1e-9 * mag) x[kept++] = v;
    }
    return kept;
}

// Solve in the better-conditioned orientation: when the leading coefficient is the smaller end
// coefficient, solve the reversed polynomial in 1/x (a near-zero leading coefficient otherwise
// sends one root to infinity and wrecks the depressed form of the others).
SLAMFE_HD int quartic_real_roots(double A4, double A3, double A2, double A1, double A0, double (&x)[4])
{
    if (fabs(A4) >= fabs(A0)) return quartic_real_roots_monic_side(A4, A3, A2, A1, A0, x);
    double y[4];
    const int n = quartic_real_roots_monic_side(A0, A1, A2, A3, A4, y);
    int kept = 0;
    for (int k = 0; k < n; ++k)
        if (y[k] != 0.0) x[kept++] = 1.0 / y[k];
    return kept;
}

struct Vec3 {
    double x, y, z;
};
SLAMFE_HD Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
SLAMFE_HD Vec3 scale(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
SLAMFE_HD double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SLAMFE_HD Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Orthonormal frame (columns e1, e2, e3) of a triangle; false if degenerate.
SLAMFE_HD bool triangle_frame(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 (&e)[3])
{
    Vec3 e1 = sub(p1, p0);
    const double n1 = sqrt(dot(e1, e1));
    Vec3 e3 = cross(e1, sub(p2, p0));
    const double n3 = sqrt(dot(e3, e3));
    if (!(n1 > 0.0) || !(n3 > 0.0)) return false;
    e1 = scale(e1, 1.0 / n1);
    e3 = scale(e3, 1.0 / n3);
    e[0] = e1;
    e[1] = cross(e3, e1);
    e[2] = e3;
    return true;
}

// Pose hypothesis from 4 correspondences.  P[i] world points, uv[i] left-image pixels (x, y),
// K (row-major 3x3) and its inverse Kinv.  T (row-major 3x4) = [R|t], world -> camera.
SLAMFE_HD bool solve_sample(const Vec3 (&P)[4], const double (&uv)[4][2], const double *K, const double *Kinv,
                            double *T)
{
    Vec3 J[3];
    for (int i = 0; i < 3; ++i) {
        const double u = uv[i][0], v = uv[i][1];
        Vec3 j = {Kinv[0] * u + Kinv[1] * v + Kinv[2], Kinv[3] * u + Kinv[4] * v + Kinv[5],
                  Kinv[6] * u + Kinv[7] * v + Kinv[8]};
        J[i] = scale(j, 1.0 / sqrt(dot(j, j)));
    }
    const Vec3 d12 = sub(P[1], P[2]), d02 = sub(P[0], P[2]), d01 = sub(P[0], P[1]);
    const double a2 = dot(d12, d12), b2 = dot(d02, d02), c2 = dot(d01, d01);
    if (!(a2 > 0.0) || !(b2 > 0.0) || !(c2 > 0.0)) return false;
    const double ca = dot(J[1], J[2]), cb = dot(J[0], J[2]), cg = dot(J[0], J[1]);
    // quartic in v = s3/s1 (sympy elimination of s1 and u = s2/s1; see oracle/p3p_oracle.py)
    const double x9 = a2 * a2 + b2 * b2 + c2 * c2 - 2.0 * a2 * c2;
    const double x15 = -ca * cg * b2 * b2 - cb * a2 * a2 - cb * c2 * c2 + 2.0 * cb * a2 * c2 + a2 * b2 * ca * cg +
                       ca * cg * b2 * c2;
    const double A4 = b2 * (-2.0 * a2 * b2 - 4.0 * b2 * c2 * ca * ca + 2.0 * b2 * c2 + x9);
    const double A3 = -4.0 * b2 * (-cb * 2.0 * b2 * c2 * ca * ca - cb * a2 * b2 + cb * b2 * c2 - x15);
    const double A2 = 2.0 * b2 *
                      (-4.0 * a2 * c2 * cb * cb - 4.0 * ca * cg * cb * a2 * b2 - 4.0 * ca * cg * cb * b2 * c2 -
                       2.0 * b2 * c2 * ca * ca + 2.0 * cg * cg * b2 * b2 - 2.0 * a2 * b2 * cg * cg +
                       2.0 * cb * cb * a2 * a2 + 2.0 * cb * cb * c2 * c2 + 2.0 * ca * ca * b2 * b2 + a2 * a2 - b2 * b2 +
                       c2 * c2 - 2.0 * a2 * c2);
    const double A1 = -4.0 * b2 * (-cb * 2.0 * a2 * b2 * cg * cg + cb * a2 * b2 - cb * b2 * c2 - x15);
    const double A0 = b2 * (2.0 * a2 * b2 - 4.0 * a2 * b2 * cg * cg - 2.0 * b2 * c2 + x9);
    double roots[4];
    const int n = quartic_real_roots(A4, A3, A2, A1, A0, roots);
    Vec3 Fw[3];
    if (!triangle_frame(P[0], P[1], P[2], Fw)) return false;
    double best_err = 1e300;
    bool found = false;
    for (int k = 0; k < n; ++k) {
        const double v = roots[k];
        if (!(v > 0.0)) continue;
        const double den = 2.0 * b2 * (cg - ca * v);
        if (fabs(den) < 1e-12 * b2) continue;
        const double w = 1.0 + v * v - 2.0 * v * cb;
        const double u = ((a2 - c2) * w - b2 * (v * v - 1.0)) / den;
        if (!(u > 0.0) || !(w > 0.0)) continue;
        const double s1 = sqrt(b2 / w);
        const Vec3 C0 = scale(J[0], s1), C1 = scale(J[1], u * s1), C2 = scale(J[2], v * s1);
        Vec3 Fc[3];
        if (!triangle_frame(C0, C1, C2, Fc)) continue;
        // R = Fc * Fw^T (columns of the frames are the basis vectors)
        double R[9];
        R[0] = Fc[0].x * Fw[0].x + Fc[1].x * Fw[1].x + Fc[2].x * Fw[2].x;
        R[1] = Fc[0].x * Fw[0].y + Fc[1].x * Fw[1].y + Fc[2].x * Fw[2].y;
        R[2] = Fc[0].x * Fw[0].z + Fc[1].x * Fw[1].z + Fc[2].x * Fw[2].z;
        R[3] = Fc[0].y * Fw[0].x + Fc[1].y * Fw[1].x + Fc[2].y * Fw[2].x;
        R[4] = Fc[0].y * Fw[0].y + Fc[1].y * Fw[1].y + Fc[2].y * Fw[2].y;
        R[5] = Fc[0].y * Fw[0].z + Fc[1].y * Fw[1].z + Fc[2].y * Fw[2].z;
        R[6] = Fc[0].z * Fw[0].x + Fc[1].z * Fw[1].x + Fc[2].z * Fw[2].x;
        R[7] = Fc[0].z * Fw[0].y + Fc[1].z * Fw[1].y + Fc[2].z * Fw[2].y;
        R[8] = Fc[0].z * Fw[0].z + Fc[1].z * Fw[1].z + Fc[2].z * Fw[2].z;
        const double tx = C0.x - (R[0] * P[0].x + R[1] * P[0].y + R[2] * P[0].z);
        const double ty = C0.y - (R[3] * P[0].x + R[4] * P[0].y + R[5] * P[0].z);
        const double tz = C0.z - (R[6] * P[0].x + R[7] * P[0].y + R[8] * P[0].z);
        // the fourth correspondence picks the solution
        const double cx = R[0] * P[3].x + R[1] * P[3].y + R[2] * P[3].z + tx;
        const double cy = R[3] * P[3].x + R[4] * P[3].y + R[5] * P[3].z + ty;
        const double cz = R[6] * P[3].x + R[7] * P[3].y + R[8] * P[3].z + tz;
        if (!(cz > 0.0)) continue;
        const double px = (K[0] * cx + K[1] * cy + K[2] * cz) / cz, py = (K[3] * cx + K[4] * cy + K[5] * cz) / cz;
        const double err = (px - uv[3][0]) * (px - uv[3][0]) + (py - uv[3][1]) * (py - uv[3][1]);
        if (err < best_err) {
            best_err = err;
            found = true;
            T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = tx;
            T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = ty;
            T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = tz;
        }
    }
    return found;
}

// Counter-based sampling: 4 distinct indices in [0, n) for (seed, frame, hypothesis) — splitmix64
// stream + Floyd's subset algorithm (every 4-subset equally likely).  n >= 4.
SLAMFE_HD uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
SLAMFE_HD void sample4(uint64_t seed, uint32_t frame, uint32_t hyp, int n, int (&idx)[4])
{
    uint64_t state = splitmix64(seed ^ (static_cast<uint64_t>(frame) << 32 | hyp));
    for (int k = 0; k < 4; ++k) {
        const int j = n - 4 + k;  // Floyd: pick t in [0, j]; if taken, take j
        state = splitmix64(state);
        int t = static_cast<int>(state % static_cast<uint64_t>(j + 1));
        for (int m = 0; m < k; ++m)
            if (idx[m] == t) t = j;
        idx[k] = t;
    }
}

}  // namespace p3p
}  // namespace slamfe
