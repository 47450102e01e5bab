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
#include "triangulate.cuh"

namespace slamfe {
namespace {

constexpr int TRI_THREADS = 256;
constexpr int TRI_ITEMS = 4;                          // links per thread
constexpr int TRI_TILE = TRI_THREADS * TRI_ITEMS;     // links per CTA

// One CTA = 1024 links.  The AoS rows [x_left, x_right, y] / [X, Y, Z] move between HBM and shared
// memory as 16-byte vectors (fully coalesced, 3 (fp32) or 6 (fp64) loads in flight per thread);
// thread t then works on links t, t+256, t+512, t+768 (stride-3 word accesses: conflict-free).
template <typename T>
__global__ void __launch_bounds__(TRI_THREADS) triangulate_links_kernel(const T *__restrict__ links, int64_t n,
                                                                        const Cams c, T *__restrict__ xyz)
{
    constexpr int VEC = 16 / sizeof(T);               // elements per 16-byte vector
    constexpr int NV = 3 * TRI_TILE / VEC;            // vectors per full tile
    __shared__ alignas(16) T s[3 * TRI_TILE];
    const int tid = threadIdx.x;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * TRI_TILE;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(TRI_TILE), n - base));
    const T *src = links + 3 * base;
    T *dst = xyz + 3 * base;
    const bool full = cnt == TRI_TILE;
    const bool vec_in = full && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const bool vec_out = full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    if (vec_in) {
        const uint4 *v = reinterpret_cast<const uint4 *>(src);
        uint4 r[NV / TRI_THREADS];
#pragma unroll
        for (int k = 0; k < NV / TRI_THREADS; ++k) r[k] = __ldcs(v + tid + k * TRI_THREADS);
#pragma unroll
        for (int k = 0; k < NV / TRI_THREADS; ++k) reinterpret_cast<uint4 *>(s)[tid + k * TRI_THREADS] = r[k];
    } else {
        for (int e = tid; e < 3 * cnt; e += TRI_THREADS) s[e] = src[e];
    }
    __syncthreads();
    T out[TRI_ITEMS][3];
#pragma unroll
    for (int k = 0; k < TRI_ITEMS; ++k) {
        const int i = tid + k * TRI_THREADS;
        double X = 0, Y = 0, Z = 0;
        if (i < cnt)
            triangulate_link(c, static_cast<double>(s[3 * i]), static_cast<double>(s[3 * i + 1]),
                             static_cast<double>(s[3 * i + 2]), X, Y, Z);
        out[k][0] = static_cast<T>(X);
        out[k][1] = static_cast<T>(Y);
        out[k][2] = static_cast<T>(Z);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TRI_ITEMS; ++k) {
        const int i = tid + k * TRI_THREADS;
        s[3 * i] = out[k][0];
        s[3 * i + 1] = out[k][1];
        s[3 * i + 2] = out[k][2];
    }
    __syncthreads();
    if (vec_out) {
#pragma unroll
        for (int k = 0; k < NV / TRI_THREADS; ++k)
            __stcs(reinterpret_cast<uint4 *>(dst) + tid + k * TRI_THREADS,
                   reinterpret_cast<const uint4 *>(s)[tid + k * TRI_THREADS]);
    } else {
        for (int e = tid; e < 3 * cnt; e += TRI_THREADS) dst[e] = s[e];
    }
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
    const int64_t blocks = (n + TRI_TILE - 1) / TRI_TILE;
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
