// TEST INFRASTRUCTURE ONLY.  Compiles the product's scorer arithmetic (csrc/ransac_core.cuh) for the
// HOST so that the CPU test suite (-m "not gpu") can fuzz the certified division-free inlier test
// against the reference formula (ransac.py:38-56 in numpy) on millions of borderline cases.
// Build with -ffp-contract=off: every fused multiply-add in the header is an explicit fma(), so the
// host build computes exactly what the device does.  The product never loads this library.
#include "../67604-slam---video-navigation_b200/csrc/ransac_core.cuh"

using namespace slamfe;

// For n points and ONE hypothesis T (3x4): out_agrees = agrees() (certified test + exact fallback),
// out_exact = agrees_exact(), out_fast = the division-free verdict alone, out_cert = 1 where the
// certificate decided (no fallback needed).
extern "C" void ransac_host_score(const double *K, const double *M1, const double *M2, const double *T,
                                  const double *pts, const double *lpix, const double *rpix, long n,
                                  unsigned char *out_agrees, unsigned char *out_exact, unsigned char *out_fast,
                                  unsigned char *out_cert, unsigned char *out_shared_rows_equal)
{
    RansacCams c;
    for (int k = 0; k < 9; ++k) c.K[k] = K[k];
    for (int k = 0; k < 12; ++k) { c.M1[k] = M1[k]; c.M2[k] = M2[k]; }
    double M[24];
    hypothesis_matrices(c, T, M, M + 12);
    const bool shared = shares_rotation_columns(M, M + 12);
    for (long i = 0; i < n; ++i) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        const double lx = lpix[2 * i], ly = lpix[2 * i + 1], rx = rpix[2 * i], ry = rpix[2 * i + 1];
        out_agrees[i] = agrees(M, x, y, z, lx, ly, rx, ry);
        out_exact[i] = agrees_exact(M, x, y, z, lx, ly, rx, ry);
        const double l0 = project_row(M + 0, x, y, z), l1 = project_row(M + 4, x, y, z), l2 = project_row(M + 8, x, y, z);
        const double r0 = project_row(M + 12, x, y, z), r1 = project_row(M + 16, x, y, z), r2 = project_row(M + 20, x, y, z);
        const RatioTest t0(l1, l2, ly, cert_of(ly)), t1(l0, l2, lx, cert_of(lx));
        const RatioTest t2(r1, r2, ry, cert_of(ry)), t3(r0, r2, rx, cert_of(rx));
        const bool out = t0.surely_outside() | t1.surely_outside() | t2.surely_outside() | t3.surely_outside();
        const bool in = t0.surely_inside() & t1.surely_inside() & t2.surely_inside() & t3.surely_inside();
        out_cert[i] = out | in;
        // the kernel's shortcut for rectified rigs: right rows = left accumulators + right 4th column
        const double a0 = project_acc(M + 0, x, y, z), a1 = project_acc(M + 4, x, y, z), a2 = project_acc(M + 8, x, y, z);
        const bool lsame = (a0 + M[3] == l0 || l0 != l0) && (a1 + M[7] == l1 || l1 != l1) && (a2 + M[11] == l2 || l2 != l2);
        const double s0 = a0 + M[15], s1 = a1 + M[19], s2 = a2 + M[23];
        const bool rsame = (s0 == r0 || r0 != r0) && (s1 == r1 || r1 != r1) && (s2 == r2 || r2 != r2);
        out_shared_rows_equal[i] = shared ? (lsame && rsame && (s0 != s0) == (r0 != r0) && (s2 != s2) == (r2 != r2)) : 2;
        out_fast[i] = (t0.diff < 0) & (t1.diff < 0) & (t2.diff < 0) & (t3.diff < 0);
    }
}

extern "C" void ransac_host_matrices(const double *K, const double *M1, const double *M2, const double *T, double *PLPR)
{
    RansacCams c;
    for (int k = 0; k < 9; ++k) c.K[k] = K[k];
    for (int k = 0; k < 12; ++k) { c.M1[k] = M1[k]; c.M2[k] = M2[k]; }
    hypothesis_matrices(c, T, PLPR, PLPR + 12);
}

// The kernel's fp32 pre-filter for n points and ONE hypothesis: out_far_v / out_far_u = the pair is dropped
// by the v / u test, out_rows = agrees_rows() with the hypothesis' own `shared` flag (what the kernel counts
// for the pairs it keeps), out_exact = the reference's verdict (agrees_exact).
extern "C" void ransac_host_prune(const double *K, const double *M1, const double *M2, const double *T,
                                  const double *pts, const double *lpix, const double *rpix, long n,
                                  unsigned char *out_far_v, unsigned char *out_far_u, unsigned char *out_rows,
                                  unsigned char *out_exact)
{
    RansacCams c;
    for (int k = 0; k < 9; ++k) c.K[k] = K[k];
    for (int k = 0; k < 12; ++k) { c.M1[k] = M1[k]; c.M2[k] = M2[k]; }
    double M[24];
    hypothesis_matrices(c, T, M, M + 12);
    const bool shared = shares_rotation_columns(M, M + 12);
    float F[12];
    for (int k = 0; k < 12; ++k) F[k] = static_cast<float>(M[k]);
    const PruneHyp ph = prune_hyp_bound_of(M);
    for (long i = 0; i < n; ++i) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        const double lx = lpix[2 * i], ly = lpix[2 * i + 1], rx = rpix[2 * i], ry = rpix[2 * i + 1];
        const float xf = static_cast<float>(x), yf = static_cast<float>(y), zf = static_cast<float>(z);
        float b, bp;
        prune_point_bound(x, y, z, lx, ly, &b, &bp);
        const float a2bp = ph.a2 * bp;
        const float l0 = prune_row(F, xf, yf, zf), l1 = prune_row(F + 4, xf, yf, zf), l2 = prune_row(F + 8, xf, yf, zf);
        out_far_v[i] = prune_far(l1, l2, static_cast<float>(ly), prune_slack(ph.a1, b, a2bp));
        out_far_u[i] = prune_far(l0, l2, static_cast<float>(lx), prune_slack(ph.a0, b, a2bp));
        out_rows[i] = agrees_rows(M, shared, x, y, z, lx, ly, rx, ry);
        out_exact[i] = agrees_exact(M, x, y, z, lx, ly, rx, ry);
    }
}
