// Arithmetic core of the RANSAC-PnP scorer (used by ransac.cu; plain C++ under the SLAMFE_HD macros so
// that oracle/ransac_host_shim.cpp can compile the very same code for the host and the CPU test suite
// can fuzz it against the reference formula — millions of borderline cases, no GPU needed).
// Every multiply-add that must be fused is an explicit fma(); nothing else relies on contraction, so
// the host build (-ffp-contract=off) and the device build produce identical bits.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "hd.cuh"

namespace slamfe {

struct RansacCams {
    double K[9];
    double M1[12];
    double M2[12];
};

// Column j of out (3x4) = (K @ T) @ [M; 0 0 0 1] for both cameras, each dot product = sequential FMA
// over k (dgemm order).  pl[i] = PL[i][j], pr[i] = PR[i][j].
SLAMFE_HD_PLAIN void hypothesis_matrix_column(const RansacCams &c, const double *T, int j, double *pl, double *pr)
{
    const double h3 = (j == 3) ? 1.0 : 0.0;  // last row of [M; 0 0 0 1]
    const double m1a = c.M1[j], m1b = c.M1[4 + j], m1c = c.M1[8 + j];
    const double m2a = c.M2[j], m2b = c.M2[4 + j], m2c = c.M2[8 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double kt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double acc = c.K[3 * i] * T[k];
            acc = fma(c.K[3 * i + 1], T[4 + k], acc);
            acc = fma(c.K[3 * i + 2], T[8 + k], acc);
            kt[k] = acc;
        }
        double a = kt[0] * m1a;
        a = fma(kt[1], m1b, a);
        a = fma(kt[2], m1c, a);
        pl[i] = fma(kt[3], h3, a);
        double b = kt[0] * m2a;
        b = fma(kt[1], m2b, b);
        b = fma(kt[2], m2c, b);
        pr[i] = fma(kt[3], h3, b);
    }
}

SLAMFE_HD_PLAIN void hypothesis_matrices(const RansacCams &c, const double *T, double *PL, double *PR)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double pl[3], pr[3];
        hypothesis_matrix_column(c, T, j, pl, pr);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            PL[4 * i + j] = pl[i];
            PR[4 * i + j] = pr[i];
        }
    }
}

// One row of a 3x4 projection times [x y z 1]: the dgemm accumulation order of numpy's matmul.
// project_acc is the part that does not involve the fourth column.
SLAMFE_HD double project_acc(const double *m, double x, double y, double z)
{
    double acc = m[0] * x;
    acc = fma(m[1], y, acc);
    acc = fma(m[2], z, acc);
    return acc;
}
SLAMFE_HD double project_row(const double *m, double x, double y, double z)
{
    return project_acc(m, x, y, z) + m[3];  // fma(m[3], 1.0, acc)
}

// Rectified rigs (M1 = [I|0], M2 = [I|t]): the left and right projection matrices of a hypothesis
// share their first three columns BIT FOR BIT, so the right camera's rows are the left camera's
// accumulators plus the right fourth column — the same bits as the full evaluation at a quarter
// of the cost.  Checked per hypothesis; NaN entries compare unequal and take the general path.
SLAMFE_HD bool shares_rotation_columns(const double *PL, const double *PR)
{
    bool same = true;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) same = same && (PL[4 * i + j] == PR[4 * i + j]);
    return same;
}

// ransac.py:38-56 for one (hypothesis, correspondence), exactly as the reference evaluates it:
// IEEE division, subtraction, strict < 2.  M = PL (12) followed by PR (12).
SLAMFE_HD_NOINLINE bool agrees_exact(const double *M, double x, double y, double z, double lx, double ly,
                                          double rx, double ry)
{
    const double l0 = project_row(M + 0, x, y, z), l1 = project_row(M + 4, x, y, z), l2 = project_row(M + 8, x, y, z);
    const double r0 = project_row(M + 12, x, y, z), r1 = project_row(M + 16, x, y, z),
                 r2 = project_row(M + 20, x, y, z);
    const double ul = l0 / l2, vl = l1 / l2, ur = r0 / r2, vr = r1 / r2;
    return (fabs(vl - ly) < 2.0) && (fabs(ul - lx) < 2.0) && (fabs(vr - ry) < 2.0) && (fabs(ur - rx) < 2.0);
}

// Division-free evaluation of  |num/den - pix| < 2  with a certificate.
// With w = num/den - pix in exact arithmetic, the reference's rounded result t = fl(fl(num/den) - pix)
// satisfies |t - w| <= (|pix| + 2|w|) * 2^-52, so its verdict equals (|w| < 2) whenever
// ||w| - 2| exceeds that.  Here e = fma(-pix, den, num) = w*den (one rounding), diff = |e| - 2|den|
// and the verdict (diff < 0) is certified when  |diff| > (|pix| + 8) * |den| * 2^-49:
//   * |e| <= 4|den|: the reference's error, scaled by |den|, is <= (|pix| + 8)|den| 2^-52 and the
//     roundings of e and diff add <= 4|den| 2^-52 — together < 1/8 of the bound;
//   * |e| >  4|den|: |diff| >= |e|/2, far above 2|e| 2^-52, and the bound covers the |pix| term.
// NaN and 0/0 never certify and take the exact path.  The fp64 divider sequences (~25
// instructions each, 4 per pair) were 3/4 of the kernel; a test now costs 3 fp64 operations.
// cert = (|pix| + 8) * 2^-49 is hoisted out of the hypothesis loop.
SLAMFE_HD double cert_of(double pix) { return (fabs(pix) + 8.0) * 0x1p-49; }

// Cheap pruning test used before the certified one: with e = fma(-pix, den, num) = w*den and ad = |den|,
//     far = (|e| - 2.001 |den| >= 0)   i.e.  |w| >= 2.001
// decided on the SIGN BIT of one FMA (an integer test: two fp64-pipe operations per coordinate instead
// of three and two compares).  The 0.1 % margin is ~10^12 times the rounding of e and of the FMA, so
// far implies the reference's |num/den - pix| < 2 is false; den = 0 gives far (reference: inf or NaN, not an
// inlier); a NaN with its sign bit set is not far and falls through to the certified / exact tests, which
// reject it as well.
SLAMFE_HD bool surely_far(double e, double ad)
{
    const double pv = fma(-2.001, ad, fabs(e));
    long long bits;
    memcpy(&bits, &pv, sizeof(bits));
    return bits >= 0;
}

struct RatioTest {
    double diff, bound;
    SLAMFE_HD RatioTest(double num, double den, double pix, double cert)
    {
        const double e = fma(-pix, den, num);
        const double ad = fabs(den);
        diff = fma(-2.0, ad, fabs(e));
        bound = cert * ad;
    }
    // from e = fma(-pix, den, num) and ad = |den| already at hand (the pruning test computed them)
    SLAMFE_HD RatioTest(double e, double ad, double cert, int)
    {
        diff = fma(-2.0, ad, fabs(e));
        bound = cert * ad;
    }
    SLAMFE_HD bool surely_outside() const { return diff > bound; }
    SLAMFE_HD bool surely_inside() const { return diff < -bound; }
};

SLAMFE_HD bool agrees(const double *M, double x, double y, double z, double lx, double ly, double rx,
                                       double ry)
{
    const double l0 = project_row(M + 0, x, y, z), l1 = project_row(M + 4, x, y, z), l2 = project_row(M + 8, x, y, z);
    const double r0 = project_row(M + 12, x, y, z), r1 = project_row(M + 16, x, y, z),
                 r2 = project_row(M + 20, x, y, z);
    const RatioTest t0(l1, l2, ly, cert_of(ly)), t1(l0, l2, lx, cert_of(lx));
    const RatioTest t2(r1, r2, ry, cert_of(ry)), t3(r0, r2, rx, cert_of(rx));
    if (t0.surely_outside() | t1.surely_outside() | t2.surely_outside() | t3.surely_outside()) return false;
    if (t0.surely_inside() & t1.surely_inside() & t2.surely_inside() & t3.surely_inside()) return true;
    return agrees_exact(M, x, y, z, lx, ly, rx, ry);
}

}  // namespace slamfe
