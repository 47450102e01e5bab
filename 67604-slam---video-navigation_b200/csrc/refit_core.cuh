// PnP refit on a consensus set: the arithmetic shared by the device kernel (refit.cu) and the host
// build the CPU tests fuzz (oracle/refit_host_shim.cpp).  Reference: the final solve of ransac_pnp,
// final_project/algorithms/ransac.py:185-193 — cv2.solvePnP(world_points, pixels, K, EPNP) on the
// left-image pixels of all inliers of the best hypothesis.  Here: Levenberg-Marquardt on the same
// objective EPnP approximates, the left-image reprojection error
//     E(R, t) = sum_i | proj(K (R X_i + t)) - pix_i |^2,
// seeded by the winning RANSAC hypothesis, pose update on the left:  R <- exp(w) R,  t <- t + dt.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SLAMFE_HD __host__ __device__ __forceinline__
#else
#define SLAMFE_HD inline
#endif

namespace slamfe {

constexpr int REFIT_NACC = 28;  // 21 (upper triangle of J^T J) + 6 (J^T r) + 1 (r.r)

// Adds one correspondence to acc[28] for the pose T (row-major 3x4) and camera K (row-major 3x3).
SLAMFE_HD void refit_accumulate(const double *T, const double *K, double X, double Y, double Z, double px, double py,
                                double *acc)
{
    // W = R X, Yc = W + t
    const double w0 = T[0] * X + T[1] * Y + T[2] * Z;
    const double w1 = T[4] * X + T[5] * Y + T[6] * Z;
    const double w2 = T[8] * X + T[9] * Y + T[10] * Z;
    const double y0 = w0 + T[3], y1 = w1 + T[7], y2 = w2 + T[11];
    const double q0 = K[0] * y0 + K[1] * y1 + K[2] * y2;
    const double q1 = K[3] * y0 + K[4] * y1 + K[5] * y2;
    const double q2 = K[6] * y0 + K[7] * y1 + K[8] * y2;
    const double iq = 1.0 / q2;
    const double u = q0 * iq, v = q1 * iq;
    const double ru = u - px, rv = v - py;
    // d(u, v) / dYc
    double a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = (K[k] - u * K[6 + k]) * iq;
        b[k] = (K[3 + k] - v * K[6 + k]) * iq;
    }
    // dYc/dw = -[W]x :  columns (0, -w2, w1), (w2, 0, -w0), (-w1, w0, 0);  dYc/dt = I
    double ju[6], jv[6];
    ju[0] = -a[1] * w2 + a[2] * w1;
    ju[1] = a[0] * w2 - a[2] * w0;
    ju[2] = -a[0] * w1 + a[1] * w0;
    ju[3] = a[0]; ju[4] = a[1]; ju[5] = a[2];
    jv[0] = -b[1] * w2 + b[2] * w1;
    jv[1] = b[0] * w2 - b[2] * w0;
    jv[2] = -b[0] * w1 + b[1] * w0;
    jv[3] = b[0]; jv[4] = b[1]; jv[5] = b[2];
    int n = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) acc[n++] += ju[r] * ju[c] + jv[r] * jv[c];
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[21 + r] += ju[r] * ru + jv[r] * rv;
    acc[27] += ru * ru + rv * rv;
}

// Solves (H + lambda diag(H)) d = -g for the 6x6 system held in acc (upper triangle, row-major) by
// Cholesky.  Returns false when the damped matrix is not positive definite.
SLAMFE_HD bool refit_solve(const double *acc, double lambda, double *d)
{
    double A[6][6];
    int n = 0;
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) {
            A[r][c] = acc[n];
            A[c][r] = acc[n];
            ++n;
        }
    for (int r = 0; r < 6; ++r) A[r][r] += lambda * A[r][r] + 1e-300;
    double L[6][6];
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = A[r][c];
            for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
            if (r == c) {
                if (!(s > 0.0)) return false;
                L[r][r] = sqrt(s);
            } else {
                L[r][c] = s / L[c][c];
            }
        }
    double y[6];
    for (int r = 0; r < 6; ++r) {
        double s = -acc[21 + r];
        for (int k = 0; k < r; ++k) s -= L[r][k] * y[k];
        y[r] = s / L[r][r];
    }
    for (int r = 5; r >= 0; --r) {
        double s = y[r];
        for (int k = r + 1; k < 6; ++k) s -= L[k][r] * d[k];
        d[r] = s / L[r][r];
    }
    return true;
}

// T <- [exp(w) R | t + dt]  with (w, dt) = d[0..5]; Rodrigues' formula in fp64.
SLAMFE_HD void refit_apply(const double *T, const double *d, double *Tn)
{
    const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const double th = sqrt(th2);
    double A, B;  // exp(w) = I + A [w]x + B [w]x^2
    if (th < 1e-8) {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
    } else {
        A = sin(th) / th;
        B = (1.0 - cos(th)) / th2;
    }
    const double wx = d[0], wy = d[1], wz = d[2];
    double E[9];
    E[0] = 1.0 - B * (wy * wy + wz * wz); E[1] = -A * wz + B * wx * wy;        E[2] = A * wy + B * wx * wz;
    E[3] = A * wz + B * wx * wy;          E[4] = 1.0 - B * (wx * wx + wz * wz); E[5] = -A * wx + B * wy * wz;
    E[6] = -A * wy + B * wx * wz;         E[7] = A * wx + B * wy * wz;          E[8] = 1.0 - B * (wx * wx + wy * wy);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Tn[4 * r + c] = E[3 * r] * T[c] + E[3 * r + 1] * T[4 + c] + E[3 * r + 2] * T[8 + c];
    Tn[3] = T[3] + d[3];
    Tn[7] = T[7] + d[4];
    Tn[11] = T[11] + d[5];
}

// One Levenberg-Marquardt controller step, run by a single thread once the sums of the CURRENT
// candidate pose are known.  State: T_acc/acc_acc/err_acc = last accepted pose with its sums, lambda.
//   returns 0 = continue (T_try holds the next pose to evaluate), 1 = converged, -1 = failed
struct RefitState {
    double T_acc[12], acc_acc[REFIT_NACC], lambda;
    int have_acc;
};

SLAMFE_HD int refit_step(RefitState &s, const double *T_cur, const double *acc_cur, double *T_try, double tol)
{
    bool accept = !s.have_acc || acc_cur[27] <= s.acc_acc[27];
    double rel = 1.0;
    if (accept) {
        if (s.have_acc) rel = (s.acc_acc[27] - acc_cur[27]) / (s.acc_acc[27] + 1e-300);
        for (int k = 0; k < 12; ++k) s.T_acc[k] = T_cur[k];
        for (int k = 0; k < REFIT_NACC; ++k) s.acc_acc[k] = acc_cur[k];
        if (s.have_acc) s.lambda = fmax(s.lambda * 0.1, 1e-12);
        s.have_acc = 1;
        if (rel < tol) return 1;
    } else {
        // a step that raises the error by a rounding-level amount means the accepted pose is the minimum
        if (acc_cur[27] - s.acc_acc[27] < tol * s.acc_acc[27]) return 1;
        s.lambda = s.lambda * 10.0;
        if (s.lambda > 1e8) return 1;  // cannot improve any further: the accepted pose is the answer
    }
    double d[6];
    int tries = 0;
    while (!refit_solve(s.acc_acc, s.lambda, d)) {
        s.lambda = fmax(s.lambda * 10.0, 1e-6);
        if (++tries > 20) return -1;
    }
    refit_apply(s.T_acc, d, T_try);
    return 0;
}

}  // namespace slamfe
