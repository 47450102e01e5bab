// RANSAC-PnP hypothesis scoring: all hypotheses x all 2D-3D correspondences, all frames, ONE
// launch — sm_100a.
//
// Replaces transformation_agreement (final_project/algorithms/ransac.py:28-56) as it is used by
// the loops ransac.py:94-112 and ransac.py:155-182: per-hypothesis inlier counts, the first
// hypothesis with the strictly largest count and its inlier mask.
//
// Bit-exactness: everything is fp64 with the reference's association order
// ((K @ T) @ [M;0001]) @ X; every dot product accumulates k = 0..3 with FMA, which is what the
// BLAS dgemm behind numpy's matmul does (checked bit-for-bit against numpy), the
// perspective division is a true IEEE division and the test is a strict < 2.  The reference's
// quirks are kept: the right camera is K @ T @ [M2;0001] (ransac.py:39) and there is no
// positive-depth test.  FP64-FMA-pipe bound (operands are reused H or N times), not HBM bound.
//
// Mapping: grid = (point tiles, hypothesis chunks, frames).  A thread owns one correspondence in
// registers and walks the chunk's hypotheses, whose 2 x 3x4 projection matrices were computed
// once per CTA into shared memory and are read with broadcast loads; votes are counted with
// ballot + popc per warp.  The last CTA of a frame (ticket counter) picks the winner and
// recomputes its mask with the same device function, so no second launch is needed.
#include "common.cuh"

namespace slamfe {
namespace {

constexpr int RS_THREADS = 256;  // correspondences per CTA
constexpr int RS_HC = 64;        // hypotheses per CTA

struct RansacCams {
    double K[9];
    double M1[12];
    double M2[12];
};

struct RansacParams {
    const double *T;
    const uint8_t *hyp_valid;
    int H;
    const double *pts, *l_pix, *r_pix;
    const int32_t *pt_off;
    int n_points;
    RansacCams cam;
    int32_t *counts, *best, *work;
    uint8_t *best_mask;
};

// out (3x4) = (K @ T) @ [M; 0 0 0 1], each dot product = sequential FMA over k (dgemm order).
__device__ void hypothesis_matrices(const RansacCams &c, const double *T, double *PL, double *PR)
{
    double KT[12];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = c.K[3 * i] * T[j];
            acc = fma(c.K[3 * i + 1], T[4 + j], acc);
            acc = fma(c.K[3 * i + 2], T[8 + j], acc);
            KT[4 * i + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double h3 = (j == 3) ? 1.0 : 0.0;  // last row of [M; 0 0 0 1]
            double a = KT[4 * i] * c.M1[j];
            a = fma(KT[4 * i + 1], c.M1[4 + j], a);
            a = fma(KT[4 * i + 2], c.M1[8 + j], a);
            a = fma(KT[4 * i + 3], h3, a);
            PL[4 * i + j] = a;
            double b = KT[4 * i] * c.M2[j];
            b = fma(KT[4 * i + 1], c.M2[4 + j], b);
            b = fma(KT[4 * i + 2], c.M2[8 + j], b);
            b = fma(KT[4 * i + 3], h3, b);
            PR[4 * i + j] = b;
        }
}

__device__ __forceinline__ double project_row(const double *m, double x, double y, double z)
{
    double acc = m[0] * x;
    acc = fma(m[1], y, acc);
    acc = fma(m[2], z, acc);
    return acc + m[3];  // fma(m[3], 1.0, acc)
}

// ransac.py:38-56 for one (hypothesis, correspondence).  M = PL (12) followed by PR (12).
__device__ __forceinline__ bool agrees(const double *M, double x, double y, double z, double lx, double ly, double rx,
                                       double ry)
{
    const double l0 = project_row(M + 0, x, y, z), l1 = project_row(M + 4, x, y, z), l2 = project_row(M + 8, x, y, z);
    const double r0 = project_row(M + 12, x, y, z), r1 = project_row(M + 16, x, y, z),
                 r2 = project_row(M + 20, x, y, z);
    const double ul = l0 / l2, vl = l1 / l2, ur = r0 / r2, vr = r1 / r2;
    return (fabs(vl - ly) < 2.0) && (fabs(ul - lx) < 2.0) && (fabs(vr - ry) < 2.0) && (fabs(ur - rx) < 2.0);
}

__global__ void __launch_bounds__(RS_THREADS) ransac_score_kernel(const RansacParams p)
{
    __shared__ alignas(16) double sM[RS_HC][24];
    __shared__ int s_cnt[RS_HC];
    __shared__ uint8_t s_valid[RS_HC];
    __shared__ unsigned long long s_red[RS_THREADS / 32];
    __shared__ int s_last, s_best;

    const int tid = threadIdx.x, lane = tid & 31;
    const int f = blockIdx.z;
    const int p0 = p.pt_off ? p.pt_off[f] : 0;
    const int np = p.pt_off ? p.pt_off[f + 1] - p0 : p.n_points;
    const int h0 = blockIdx.y * RS_HC;
    const int nh = min(RS_HC, p.H - h0);
    const int i = blockIdx.x * RS_THREADS + tid;
    const size_t hbase = static_cast<size_t>(f) * p.H;

    if (blockIdx.x * RS_THREADS < np) {  // CTA-uniform
        if (tid < RS_HC) {
            s_cnt[tid] = 0;
            bool ok = tid < nh;
            if (ok && p.hyp_valid) ok = p.hyp_valid[hbase + h0 + tid] != 0;
            s_valid[tid] = ok;
            if (ok) hypothesis_matrices(p.cam, p.T + (hbase + h0 + tid) * 12, &sM[tid][0], &sM[tid][12]);
        }
        double x = 0, y = 0, z = 0, lx = 0, ly = 0, rx = 0, ry = 0;
        const bool have = i < np;
        if (have) {
            const size_t g = static_cast<size_t>(p0 + i);
            x = p.pts[3 * g]; y = p.pts[3 * g + 1]; z = p.pts[3 * g + 2];
            lx = p.l_pix[2 * g]; ly = p.l_pix[2 * g + 1];
            rx = p.r_pix[2 * g]; ry = p.r_pix[2 * g + 1];
        }
        __syncthreads();
        for (int h = 0; h < nh; ++h) {
            if (!s_valid[h]) continue;  // CTA-uniform
            const bool in = have && agrees(&sM[h][0], x, y, z, lx, ly, rx, ry);
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, in);
            if (lane == 0 && bal) atomicAdd(&s_cnt[h], __popc(bal));
        }
        __syncthreads();
        if (tid < nh && s_cnt[tid]) atomicAdd(p.counts + hbase + h0 + tid, s_cnt[tid]);
    }

    // ---- last CTA of this frame: winner + its mask ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(p.work + f, 1);
        s_last = (ticket == static_cast<int>(gridDim.x * gridDim.y) - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // argmax over counts, lowest index on ties (ransac.py:110 keeps strictly better only)
    unsigned long long bestk = 0;
    for (int h = tid; h < p.H; h += RS_THREADS) {
        const int c = __ldcg(p.counts + hbase + h);
        const unsigned long long k =
            (static_cast<unsigned long long>(static_cast<uint32_t>(c)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(h));
        bestk = max(bestk, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bestk = max(bestk, __shfl_xor_sync(0xFFFFFFFFu, bestk, o));
    if (lane == 0) s_red[tid >> 5] = bestk;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) b = max(b, s_red[w]);
        const int cnt = static_cast<int>(b >> 32);
        const int idx = cnt > 0 ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(b)) : -1;
        p.best[2 * f] = idx;
        p.best[2 * f + 1] = cnt;
        s_best = idx;
        if (idx >= 0) hypothesis_matrices(p.cam, p.T + (hbase + idx) * 12, &sM[0][0], &sM[0][12]);
    }
    __syncthreads();
    const int bi = s_best;
    for (int k = tid; k < np; k += RS_THREADS) {
        uint8_t m = 0;
        if (bi >= 0) {
            const size_t g = static_cast<size_t>(p0 + k);
            m = agrees(&sM[0][0], p.pts[3 * g], p.pts[3 * g + 1], p.pts[3 * g + 2], p.l_pix[2 * g], p.l_pix[2 * g + 1],
                       p.r_pix[2 * g], p.r_pix[2 * g + 1])
                    ? 1
                    : 0;
        }
        p.best_mask[p0 + k] = m;
    }
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_ransac_score(const double *T, const uint8_t *hyp_valid, int H, const double *pts,
                                   const double *l_pix, const double *r_pix, const int32_t *pt_off, int n_points,
                                   int n_frames, int max_points, const double *K, const double *M1, const double *M2,
                                   int32_t *counts, int32_t *best, uint8_t *best_mask, int32_t *work,
                                   slamfe_stream_t stream)
{
    if (H < 0 || n_frames < 0 || max_points < 0 || n_points < 0) return SLAMFE_EINVAL;
    if (n_frames == 0) return 0;
    if (!best || !work || !K || !M1 || !M2) return SLAMFE_EINVAL;
    if (H > 0 && (!T || !counts)) return SLAMFE_EINVAL;
    if (!pt_off && n_frames != 1) return SLAMFE_EINVAL;
    if (!pt_off) max_points = n_points;
    if (max_points > 0 && (!pts || !l_pix || !r_pix || !best_mask)) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SLAMFE_CUDA_OK(cudaMemsetAsync(work, 0, sizeof(int32_t) * n_frames, s));
    if (H > 0) SLAMFE_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * static_cast<size_t>(n_frames) * H, s));
    RansacParams p{};
    p.T = T; p.hyp_valid = hyp_valid; p.H = H;
    p.pts = pts; p.l_pix = l_pix; p.r_pix = r_pix; p.pt_off = pt_off; p.n_points = n_points;
    for (int k = 0; k < 9; ++k) p.cam.K[k] = K[k];
    for (int k = 0; k < 12; ++k) { p.cam.M1[k] = M1[k]; p.cam.M2[k] = M2[k]; }
    p.counts = counts; p.best = best; p.work = work; p.best_mask = best_mask;
    const dim3 grid(max(1, (max_points + RS_THREADS - 1) / RS_THREADS), max(1, (H + RS_HC - 1) / RS_HC), n_frames);
    if (grid.y > 65535u || grid.z > 65535u) return SLAMFE_ERANGE;
    ransac_score_kernel<<<grid, RS_THREADS, 0, s>>>(p);
    return launch_status();
}
