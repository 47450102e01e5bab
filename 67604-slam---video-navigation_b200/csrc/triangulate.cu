// Batched stereo DLT triangulation — sm_100a.
//
// Replaces linear_least_squares_triangulation / triangulate_links / triangulate_last_frame
// (final_project/algorithms/triangulation.py:5-50): one thread per match, everything in fp64
// registers, AoS rows staged through shared memory so that global loads / stores are coalesced.
//
//  * links path (x_left, y), (x_right, y) with P[1]==Q[1], P[2]==Q[2]: rows 1 and 3 of the DLT
//    matrix (triangulation.py:18,20) are identical, the matrix has rank 3 and its null vector is
//    the cofactor vector of the three distinct rows — exact, no iteration: HBM-bound
//    (24 B per match in fp32, 48 B in fp64).
//  * general path (arbitrary P, Q, distinct y — analysis.py:400, VAN_ex/code/ex2.py:209):
//    smallest right singular vector of the 4x4 matrix by one-sided (Hestenes) Jacobi.
#include "common.cuh"

namespace slamfe {
namespace {

struct Cams {
    double P[12];
    double Q[12];
};

__device__ __forceinline__ double det3(double a0, double a1, double a2, double b0, double b1, double b2, double c0,
                                       double c1, double c2)
{
    return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
}

// Null vector of the rank-3 DLT system of a link; returns X = n[:3] / n[3].
__device__ __forceinline__ void triangulate_link(const Cams &c, double xl, double xr, double y, double &X, double &Y,
                                                 double &Z)
{
    double a[4], b[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = c.P[8 + k] * xl - c.P[k];      // triangulation.py:17
        b[k] = c.P[8 + k] * y - c.P[4 + k];   // triangulation.py:18 (== :20 for links)
        d[k] = c.Q[8 + k] * xr - c.Q[k];      // triangulation.py:19
    }
    const double n0 = det3(a[1], a[2], a[3], b[1], b[2], b[3], d[1], d[2], d[3]);
    const double n1 = -det3(a[0], a[2], a[3], b[0], b[2], b[3], d[0], d[2], d[3]);
    const double n2 = det3(a[0], a[1], a[3], b[0], b[1], b[3], d[0], d[1], d[3]);
    double n3 = -det3(a[0], a[1], a[2], b[0], b[1], b[2], d[0], d[1], d[2]);
    double s = 1.0;
    if (n3 == 0.0) {  // triangulation.py:22-23 guard on the unit-norm singular vector
        s = rsqrt(n0 * n0 + n1 * n1 + n2 * n2);
        n3 = 1e-20;
    }
    X = n0 * s / n3;
    Y = n1 * s / n3;
    Z = n2 * s / n3;
}

constexpr int TRI_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(TRI_THREADS) triangulate_links_kernel(const T *__restrict__ links, int64_t n,
                                                                        const Cams c, T *__restrict__ xyz)
{
    __shared__ T s[3 * TRI_THREADS];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * TRI_THREADS;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(TRI_THREADS), n - base));
    for (int e = threadIdx.x; e < 3 * cnt; e += TRI_THREADS) s[e] = links[3 * base + e];
    __syncthreads();
    double X = 0, Y = 0, Z = 0;
    if (threadIdx.x < cnt)
        triangulate_link(c, static_cast<double>(s[3 * threadIdx.x]), static_cast<double>(s[3 * threadIdx.x + 1]),
                         static_cast<double>(s[3 * threadIdx.x + 2]), X, Y, Z);
    __syncthreads();
    if (threadIdx.x < cnt) {
        s[3 * threadIdx.x] = static_cast<T>(X);
        s[3 * threadIdx.x + 1] = static_cast<T>(Y);
        s[3 * threadIdx.x + 2] = static_cast<T>(Z);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 3 * cnt; e += TRI_THREADS) xyz[3 * base + e] = s[e];
}

// One Jacobi rotation between columns p and q of G (4x4, column norms -> singular values).
#define SLAMFE_JACOBI_PAIR(p, q)                                                                     \
    {                                                                                                \
        double alpha = 0, beta = 0, gamma = 0;                                                       \
        _Pragma("unroll") for (int i = 0; i < 4; ++i)                                                \
        {                                                                                            \
            alpha = fma(G[i][p], G[i][p], alpha);                                                    \
            beta = fma(G[i][q], G[i][q], beta);                                                      \
            gamma = fma(G[i][p], G[i][q], gamma);                                                    \
        }                                                                                            \
        if (gamma != 0.0 && fabs(gamma) > 1e-16 * sqrt(alpha * beta)) {                              \
            rotated = true;                                                                          \
            const double zeta = (beta - alpha) / (2.0 * gamma);                                      \
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));           \
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;                                  \
            _Pragma("unroll") for (int i = 0; i < 4; ++i)                                            \
            {                                                                                        \
                const double gp = G[i][p], gq = G[i][q];                                             \
                G[i][p] = cs * gp - sn * gq;                                                         \
                G[i][q] = sn * gp + cs * gq;                                                         \
                const double vp = V[i][p], vq = V[i][q];                                             \
                V[i][p] = cs * vp - sn * vq;                                                         \
                V[i][q] = sn * vp + cs * vq;                                                         \
            }                                                                                        \
        }                                                                                            \
    }

__global__ void __launch_bounds__(128) triangulate_dlt_kernel(const double2 *__restrict__ pxy,
                                                              const double2 *__restrict__ qxy, int64_t n, const Cams c,
                                                              double *__restrict__ xyz)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = pxy[i], q = qxy[i];
    double G[4][4], V[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        G[0][k] = c.P[8 + k] * p.x - c.P[k];      // triangulation.py:17
        G[1][k] = c.P[8 + k] * p.y - c.P[4 + k];  // triangulation.py:18
        G[2][k] = c.Q[8 + k] * q.x - c.Q[k];      // triangulation.py:19
        G[3][k] = c.Q[8 + k] * q.y - c.Q[4 + k];  // triangulation.py:20
#pragma unroll
        for (int r = 0; r < 4; ++r) V[r][k] = (r == k) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        SLAMFE_JACOBI_PAIR(0, 1)
        SLAMFE_JACOBI_PAIR(0, 2)
        SLAMFE_JACOBI_PAIR(0, 3)
        SLAMFE_JACOBI_PAIR(1, 2)
        SLAMFE_JACOBI_PAIR(1, 3)
        SLAMFE_JACOBI_PAIR(2, 3)
        if (!rotated) break;
    }
    // smallest singular value = smallest column norm; its V column is the null direction
    double best = 0, v0 = 0, v1 = 0, v2 = 0, v3 = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double nn = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) nn = fma(G[r][k], G[r][k], nn);
        if (k == 0 || nn < best) {
            best = nn;
            v0 = V[0][k];
            v1 = V[1][k];
            v2 = V[2][k];
            v3 = V[3][k];
        }
    }
    if (v3 == 0.0) v3 = 1e-20;  // triangulation.py:22-23
    xyz[3 * i + 0] = v0 / v3;
    xyz[3 * i + 1] = v1 / v3;
    xyz[3 * i + 2] = v2 / v3;
}

int load_cams(const double *P, const double *Q, Cams &c)
{
    if (!P || !Q) return SLAMFE_EINVAL;
    for (int k = 0; k < 12; ++k) {
        c.P[k] = P[k];
        c.Q[k] = Q[k];
    }
    return 0;
}

template <typename T>
int run_links(const T *links, int64_t n, const double *P, const double *Q, T *xyz, cudaStream_t stream)
{
    if (n < 0) return SLAMFE_EINVAL;
    if (n == 0) return 0;
    if (!links || !xyz) return SLAMFE_EINVAL;
    Cams c;
    const int rc = load_cams(P, Q, c);
    if (rc) return rc;
    for (int k = 4; k < 12; ++k)
        if (c.P[k] != c.Q[k]) return SLAMFE_EINVAL;  // not a shared-row stereo pair: use the DLT entry point
    const int64_t blocks = (n + TRI_THREADS - 1) / TRI_THREADS;
    if (blocks > 0x7FFFFFFF) return SLAMFE_ERANGE;
    triangulate_links_kernel<T><<<static_cast<unsigned>(blocks), TRI_THREADS, 0, stream>>>(links, n, c, xyz);
    return launch_status();
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_triangulate_links_f64(const double *links, int64_t n, const double *P, const double *Q,
                                            double *xyz, slamfe_stream_t stream)
{
    return run_links<double>(links, n, P, Q, xyz, static_cast<cudaStream_t>(stream));
}

extern "C" int slamfe_triangulate_links_f32(const float *links, int64_t n, const double *P, const double *Q, float *xyz,
                                            slamfe_stream_t stream)
{
    return run_links<float>(links, n, P, Q, xyz, static_cast<cudaStream_t>(stream));
}

extern "C" int slamfe_triangulate_dlt_f64(const double *pxy, const double *qxy, int64_t n, const double *P,
                                          const double *Q, double *xyz, slamfe_stream_t stream)
{
    if (n < 0) return SLAMFE_EINVAL;
    if (n == 0) return 0;
    if (!pxy || !qxy || !xyz) return SLAMFE_EINVAL;
    Cams c;
    const int rc = load_cams(P, Q, c);
    if (rc) return rc;
    const int64_t blocks = (n + 127) / 128;
    if (blocks > 0x7FFFFFFF) return SLAMFE_ERANGE;
    triangulate_dlt_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double2 *>(pxy), reinterpret_cast<const double2 *>(qxy), n, c, xyz);
    return launch_status();
}
