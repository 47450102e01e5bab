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

struct RatioTest {
    double diff, bound;
    SLAMFE_HD RatioTest(double num, double den, double pix, double cert)
    {
        const double e = fma(-pix, den, num);
        const double ad = fabs(den);
        diff = fma(-2.0, ad, fabs(e));
        bound = cert * ad;
    }
    SLAMFE_HD bool surely_outside() const { return diff > bound; }
    SLAMFE_HD bool surely_inside() const { return diff < -bound; }
};

// The scorer's evaluation of one (hypothesis, correspondence): certified division-free tests, exact
// fallback.  `shared` = shares_rotation_columns(M, M + 12): the right rows are then the left accumulators
// plus the right fourth column (the same bits).  Left camera first: its verdict alone rejects most pairs.
SLAMFE_HD bool agrees_rows(const double *M, bool shared, double x, double y, double z, double lx, double ly,
                           double rx, double ry)
{
    const double a0 = project_acc(M + 0, x, y, z), a1 = project_acc(M + 4, x, y, z), a2 = project_acc(M + 8, x, y, z);
    const double l2 = a2 + M[11];
    const RatioTest t0(a1 + M[7], l2, ly, cert_of(ly)), t1(a0 + M[3], l2, lx, cert_of(lx));
    if (t0.surely_outside() | t1.surely_outside()) return false;
    double r0, r1, r2;
    if (shared) {
        r0 = a0 + M[15]; r1 = a1 + M[19]; r2 = a2 + M[23];
    } else {
        r0 = project_row(M + 12, x, y, z); r1 = project_row(M + 16, x, y, z); r2 = project_row(M + 20, x, y, z);
    }
    const RatioTest t2(r1, r2, ry, cert_of(ry)), t3(r0, r2, rx, cert_of(rx));
    if (t2.surely_outside() | t3.surely_outside()) return false;
    if (t0.surely_inside() & t1.surely_inside() & t2.surely_inside() & t3.surely_inside()) return true;
    return agrees_exact(M, x, y, z, lx, ly, rx, ry);
}

SLAMFE_HD bool agrees(const double *M, double x, double y, double z, double lx, double ly, double rx,
                                       double ry)
{
    return agrees_rows(M, false, x, y, z, lx, ly, rx, ry);
}

// ---- fp32 pre-filter -------------------------------------------------------------------------------
// Most (hypothesis, correspondence) pairs of a RANSAC run are nowhere near agreement.  Before any fp64
// work the scorer evaluates the LEFT camera in fp32 and drops a pair when that alone proves the
// reference's verdict false; a warp whose pairs are all dropped does nothing else for the hypothesis.
// The filter is one-sided: it may keep a pair (which then gets the full fp64 evaluation above), it never
// drops one the reference accepts.
//
// Notation: m_i = row i of the left matrix PL (fp64, as the scorer holds it), X = (x, y, z, 1),
// S_i = sum_j |m_ij X_j|, L_i = the row value the reference computes (fp64), u = 2^-24.
//   * l~_i = prune_row(fp32(m_i), fp32(X)) is three FMAs on rounded inputs: every term passes through at
//     most five factors (1 + d), |d| <= u, so |l~_i - L_i| <= 5.01 u S_i + (the reference's own 4 * 2^-53 S_i)
//     < 2^-20 S_i.
//   * e~ = fmaf(-fp32(pix), l~_2, l~_1) is one more rounding:  |e~ - (L_1 - pix L_2)| <= 1.13 * 2^-20 (S_1 + |pix| S_2).
//   * S_i <= A_i B  with  A_i = max_j |m_ij|  and  B = |x| + |y| + |z| + 1.
// The v test is  fmaf(-2.001f, |l~_2|, |e~|) >= slack_v,  slack_v = 2^-18 B (A_1 + (p + 2) A_2),  p = max(|lx|, |ly|)
// (u: A_0 for A_1), evaluated in fp32 from factors that were each rounded UP by (1 + 2^-20).  The FMA of the
// test rounds once, relative to its own result, so a true comparison means, in exact arithmetic,
//     |e~| >= 2.0009 |l~_2| + 0.9999 slack_v,
// and with the error bounds above (slack_v is at least 1.99 times what they need)
//     |L_1 - pix L_2| >= 2.0009 |l~_2| + 2.01 * 2^-20 A_2 B >= (2 + 10^-6) |L_2|,
// i.e. |L_1/L_2 - pix| >= 2 + 10^-6 exactly, which the reference's two roundings (division, subtraction:
// <= 2^-52 (|L_1/L_2| + |pix|)) cannot bring below 2 for |pix| <= 10^9; L_2 = 0 gives inf or NaN, not an
// inlier either.  Range guards make the rest of IEEE harmless: the factors are +inf (the test is then never
// true) unless every |entry of PL|, B, |lx|, |ly| <= 10^9, which rules out fp32 overflow (|e~| < 10^28) and
// NaN inputs; a 2^-60 floor on each A_i keeps the slack (>= 2^-78) far above anything fp32 underflow can
// lose (<= 10^9 * 2^-140 with denormals, which this build keeps: no -ftz).  The comparison is an ordinary >=,
// false on NaN.
struct PruneHyp {
    float a1, a2, a0;  // rounded-up row bounds A_1, A_2, A_0 (this order: the v test comes first)
};
SLAMFE_HD float prune_round_up(double v) { return static_cast<float>(v) * (1.0f + 0x1p-20f); }
// amax[i] = max_j |PL[i][j]|; finite = every |entry| <= 1e9 (false if any is NaN)
SLAMFE_HD PruneHyp prune_hyp_bound(const double *amax, bool finite)
{
    PruneHyp h;
    if (!finite) {
        h.a0 = h.a1 = h.a2 = INFINITY;
    } else {
        h.a0 = prune_round_up(fmax(amax[0], 0x1p-60));
        h.a1 = prune_round_up(fmax(amax[1], 0x1p-60));
        h.a2 = prune_round_up(fmax(amax[2], 0x1p-60));
    }
    return h;
}
SLAMFE_HD PruneHyp prune_hyp_bound_of(const double *PL)
{
    double amax[3] = {0.0, 0.0, 0.0};
    bool finite = true;
    for (int i = 0; i < 12; ++i) {
        finite = finite && (fabs(PL[i]) <= 1e9);  // false on NaN
        amax[i / 4] = fmax(amax[i / 4], fabs(PL[i]));
    }
    return prune_hyp_bound(amax, finite);
}
// b = 2^-18 B and bp = 2^-18 B (p + 2), rounded up; +inf outside the guarded range
SLAMFE_HD void prune_point_bound(double x, double y, double z, double lx, double ly, float *b, float *bp)
{
    const double bb = fabs(x) + fabs(y) + fabs(z) + 1.0;
    if (!(bb <= 1e9) || !(fabs(lx) <= 1e9) || !(fabs(ly) <= 1e9)) {  // NaN lands here too
        *b = *bp = INFINITY;
        return;
    }
    *b = prune_round_up(0x1p-18 * bb);
    *bp = prune_round_up(0x1p-18 * bb * (fmax(fabs(lx), fabs(ly)) + 2.0));
}
// slack of one test: a = A_1 (v) or A_0 (u), a2bp = A_2 * bp (shared by the two tests)
SLAMFE_HD float prune_slack(float a, float b, float a2bp) { return fmaf(a, b, a2bp); }
SLAMFE_HD float prune_row(const float *m, float x, float y, float z)
{
    return fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
}
SLAMFE_HD bool prune_far(float num, float den, float pix, float slack)
{
    const float e = fmaf(-pix, den, num);
    return fmaf(-2.001f, fabsf(den), fabsf(e)) >= slack;
}

}  // namespace slamfe
