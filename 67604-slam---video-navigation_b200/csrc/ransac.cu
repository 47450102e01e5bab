// RANSAC-PnP hypothesis scoring: all hypotheses x all 2D-3D correspondences, all frames, ONE
// launch — sm_100a.
//
// Replaces transformation_agreement (final_project/algorithms/ransac.py:28-56) as it is used by
// the loops ransac.py:94-112 and ransac.py:155-182: per-hypothesis inlier counts, the first
// hypothesis with the strictly largest count and its inlier mask.
//
// Bit-exactness: everything is fp64 with the reference's association order
// ((K @ T) @ [M;0001]) @ X; every dot product accumulates k = 0..3 with FMA, which is what the
// BLAS dgemm behind numpy's matmul does (checked bit-for-bit against numpy).  The
// reference's "IEEE division, subtract, strict < 2" verdict is reproduced without dividing:
// a multiplication-form test with a rounding certificate decides all but the (practically
// never occurring) borderline cases, which fall back to the literal division (agrees / agrees_exact).  The reference's
// quirks are kept: the right camera is K @ T @ [M2;0001] (ransac.py:39) and there is no
// positive-depth test.  FP64-FMA-pipe bound (operands are reused H or N times), not HBM bound.
//
// Mapping: grid = (point tiles, hypothesis splits, frames).  A thread owns RS_PP correspondences in
// registers for the CTA's whole life and walks the hypotheses of chunk blockIdx.y, blockIdx.y + gridDim.y, ...
// (RS_HC per chunk); a chunk's 2 x 3x4 projection matrices are built by all 256 threads (4 per
// hypothesis, one matrix column each) into one half of a double-buffered shared array — one barrier per
// chunk — and read back with broadcast loads; votes are counted with ballot + popc per warp.  The host
// picks the number of splits so that the grid still fills the GPU several times over: with thousands of
// frames (loop-closure candidates) one CTA sweeps all hypotheses and its correspondence loads and
// start-up are paid once, not once per 64 hypotheses.  The last CTA of a frame (ticket counter) picks the
// winner and recomputes its mask with the same device function, so no second launch is needed.
#include "common.cuh"
#include "ransac_core.cuh"

namespace slamfe {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_PP = 2;                        // correspondences per thread
constexpr int RS_TILE = RS_THREADS * RS_PP;     // correspondences per CTA
constexpr int RS_HC = 64;                       // hypotheses per CTA

struct RansacParams {
    const double *T;
    const uint8_t *hyp_valid;
    int H;
    const double *pts, *l_pix, *r_pix;
    const int32_t *pt_off, *pt_cnt;
    int n_points;
    RansacCams cam;
    int32_t *counts, *best, *work;
    uint8_t *best_mask;
};

__global__ void __launch_bounds__(RS_THREADS, 4) ransac_score_kernel(const RansacParams p)
{
    __shared__ alignas(16) double sM[2][RS_HC][24];
    __shared__ alignas(16) float sF[2][RS_HC][16];  // fp32 pre-filter: PL rows 1, 2, 0 (4 each), A_1, A_2, A_0, unused
    __shared__ int s_cnt[3][RS_HC];  // three deep: a slot is flushed one barrier after its chunk and zeroed two after
    __shared__ uint8_t s_shared[2][RS_HC];  // hypothesis' left/right matrices share their first three columns
    __shared__ unsigned long long s_red[RS_THREADS / 32];
    __shared__ int s_last, s_best;
    static_assert(RS_HC == 64 && RS_THREADS == 4 * RS_HC, "one 64-bit mask per chunk, 4 threads per hypothesis");

    const int tid = threadIdx.x, lane = tid & 31;
    const int f = blockIdx.z;
    const int p0 = p.pt_off ? p.pt_off[f] : 0;
    const int np = p.pt_cnt ? p.pt_cnt[f] : (p.pt_off ? p.pt_off[f + 1] - p0 : p.n_points);
    const size_t hbase = static_cast<size_t>(f) * p.H;
    const int n_chunks = (p.H + RS_HC - 1) / RS_HC;

    if (blockIdx.x * RS_TILE < np) {  // CTA-uniform
        // RS_PP correspondences per thread.  The sweep keeps only the fp32 pre-filter's view of them in registers
        // (coordinates, left pixel, the two slack factors); the fp64 originals are re-read from global memory
        // (L1) by the one trip in ten that gets past the filter.
        float xf[RS_PP], yf[RS_PP], zf[RS_PP], lxf[RS_PP], lyf[RS_PP], pb[RS_PP], pbp[RS_PP];
        bool have[RS_PP];
#pragma unroll
        for (int k = 0; k < RS_PP; ++k) {
            const int i = blockIdx.x * RS_TILE + k * RS_THREADS + tid;
            have[k] = i < np;
            double x = 0.0, y = 0.0, z = 0.0, lx = 0.0, ly = 0.0;
            if (have[k]) {
                const size_t g = static_cast<size_t>(p0 + i);
                x = p.pts[3 * g]; y = p.pts[3 * g + 1]; z = p.pts[3 * g + 2];
                lx = p.l_pix[2 * g]; ly = p.l_pix[2 * g + 1];
            }
            xf[k] = static_cast<float>(x); yf[k] = static_cast<float>(y); zf[k] = static_cast<float>(z);
            lxf[k] = static_cast<float>(lx); lyf[k] = static_cast<float>(ly);
            prune_point_bound(x, y, z, lx, ly, &pb[k], &pbp[k]);
            // a slot past the end of the frame's correspondences: slack -inf, so the filter always drops it
            // (the slack is a sum of products of these with positive factors, never inf - inf)
            if (!have[k]) pb[k] = pbp[k] = -INFINITY;
        }
        // flush of the previous chunk's votes: slot and hypothesis of this thread's (j == 0 lanes only)
        int prev_slot = -1, prev_hyp = 0, it = 0;
        for (int c = blockIdx.y; c < n_chunks; c += gridDim.y, ++it) {
            const int buf = it & 1, cb = it % 3;
            const int h0 = c * RS_HC;
            const int nh = min(RS_HC, p.H - h0);
            // Valid hypotheses of the chunk, as a 64-bit mask every warp computes for itself (no barrier): the
            // matrices go to shared memory COMPACTED, in order, so the sweep below has no per-hypothesis
            // validity load or branch and makes no trip for an invalid one (a quarter of the P3P hypotheses).
            unsigned long long valid;
            {
                bool v0 = lane < nh, v1 = lane + 32 < nh;
                if (p.hyp_valid) {
                    v0 = v0 && p.hyp_valid[hbase + h0 + lane] != 0;
                    v1 = v1 && p.hyp_valid[hbase + h0 + 32 + lane] != 0;
                }
                valid = static_cast<unsigned long long>(__ballot_sync(0xFFFFFFFFu, v0)) |
                        (static_cast<unsigned long long>(__ballot_sync(0xFFFFFFFFu, v1)) << 32);
            }
            const int n_valid = __popcll(valid);
            const int hh = tid >> 2, j = tid & 3;  // thread (hh, j) builds column j of hypothesis hh's PL and PR
            const bool ok = (valid >> hh) & 1;
            const int slot = __popcll(valid & ((1ull << hh) - 1ull));
            if (ok) {
                double pl[3], pr[3];
                hypothesis_matrix_column(p.cam, p.T + (hbase + h0 + hh) * 12, j, pl, pr);
                bool same = true, finite = true;
                double amax[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    sM[buf][slot][4 * i + j] = pl[i];
                    sM[buf][slot][12 + 4 * i + j] = pr[i];
                    same = same && (j == 3 || pl[i] == pr[i]);  // shares_rotation_columns, this column's part
                    finite = finite && (fabs(pl[i]) <= 1e9);
                    amax[i] = fabs(pl[i]);
                }
                sF[buf][slot][j] = static_cast<float>(pl[1]);
                sF[buf][slot][4 + j] = static_cast<float>(pl[2]);
                sF[buf][slot][8 + j] = static_cast<float>(pl[0]);
                // the 4 lanes of a hypothesis are converged here (ok is the same for all of them)
                const uint32_t group = 0xFu << (lane & ~3);
                const uint32_t sameb = __ballot_sync(group, same);
                const uint32_t finb = __ballot_sync(group, finite);
#pragma unroll
                for (int i = 0; i < 3; ++i) {  // row maxima over the 4 columns
                    amax[i] = fmax(amax[i], __shfl_xor_sync(group, amax[i], 1));
                    amax[i] = fmax(amax[i], __shfl_xor_sync(group, amax[i], 2));
                }
                if (j == 0) {
                    s_cnt[cb][slot] = 0;
                    s_shared[buf][slot] = (sameb & group) == group;
                    const PruneHyp ph = prune_hyp_bound(amax, (finb & group) == group);
                    sF[buf][slot][12] = ph.a1; sF[buf][slot][13] = ph.a2; sF[buf][slot][14] = ph.a0;
                }
            }
            __syncthreads();  // the matrices are visible, and every warp is done with the previous chunk
            if (prev_slot >= 0 && s_cnt[(it + 2) % 3][prev_slot])
                atomicAdd(p.counts + prev_hyp, s_cnt[(it + 2) % 3][prev_slot]);
            prev_slot = (ok && j == 0) ? slot : -1;
            prev_hyp = static_cast<int>(hbase) + h0 + hh;
#pragma unroll 2
            for (int h = 0; h < n_valid; ++h) {
                const float *F = &sF[buf][h][0];
                // (a) fp32 pre-filter (ransac_core.cuh) on the left image's v: rows 1 and 2.  A bad hypothesis
                // puts v more than 2 px off for every point of the warp half of the time, and then nothing else
                // is computed for it.
                const float a1 = F[12], a2 = F[13];
                float l2f[RS_PP], a2bp[RS_PP];
                bool maybe[RS_PP], any = false;
#pragma unroll
                for (int k = 0; k < RS_PP; ++k) {
                    const float l1f = prune_row(F, xf[k], yf[k], zf[k]);
                    l2f[k] = prune_row(F + 4, xf[k], yf[k], zf[k]);
                    a2bp[k] = a2 * pbp[k];
                    maybe[k] = !prune_far(l1f, l2f[k], lyf[k], prune_slack(a1, pb[k], a2bp[k]));
                    any |= maybe[k];
                }
                if (!__any_sync(0xFFFFFFFFu, any)) continue;  // warp-uniform
                // (b) the same on u: row 0
                any = false;
                const float a0 = F[14];
#pragma unroll
                for (int k = 0; k < RS_PP; ++k) {
                    const float l0f = prune_row(F + 8, xf[k], yf[k], zf[k]);
                    maybe[k] = maybe[k] && !prune_far(l0f, l2f[k], lxf[k], prune_slack(a0, pb[k], a2bp[k]));
                    any |= maybe[k];
                }
                if (!__any_sync(0xFFFFFFFFu, any)) continue;  // warp-uniform
                // (c) fp64, bit-exact (about one trip in ten gets here): certified tests, exact fallback
                const double *M = &sM[buf][h][0];
                const bool shared = s_shared[buf][h];  // CTA-uniform
#pragma unroll
                for (int k = 0; k < RS_PP; ++k) {
                    bool in = false;
                    const int i = blockIdx.x * RS_TILE + k * RS_THREADS + tid;
                    // (an empty slot is dropped by its -inf slack unless the hypothesis holds a NaN: checked here)
                    if (maybe[k] && i < np) {
                        const size_t g = static_cast<size_t>(p0 + i);
                        in = agrees_rows(M, shared, __ldg(p.pts + 3 * g), __ldg(p.pts + 3 * g + 1), __ldg(p.pts + 3 * g + 2),
                                         __ldg(p.l_pix + 2 * g), __ldg(p.l_pix + 2 * g + 1), __ldg(p.r_pix + 2 * g),
                                         __ldg(p.r_pix + 2 * g + 1));
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, in);
                    if (lane == 0 && bal) atomicAdd(&s_cnt[cb][h], __popc(bal));
                }
            }
        }
        __syncthreads();
        if (prev_slot >= 0 && s_cnt[(it + 2) % 3][prev_slot]) atomicAdd(p.counts + prev_hyp, s_cnt[(it + 2) % 3][prev_slot]);
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
        if (idx >= 0) hypothesis_matrices(p.cam, p.T + (hbase + idx) * 12, &sM[0][0][0], &sM[0][0][12]);
    }
    __syncthreads();
    const int bi = s_best;
    for (int k = tid; k < np; k += RS_THREADS) {
        uint8_t m = 0;
        if (bi >= 0) {
            const size_t g = static_cast<size_t>(p0 + k);
            m = agrees(&sM[0][0][0], p.pts[3 * g], p.pts[3 * g + 1], p.pts[3 * g + 2], p.l_pix[2 * g], p.l_pix[2 * g + 1],
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
                                   const double *l_pix, const double *r_pix, const int32_t *pt_off,
                                   const int32_t *pt_cnt, int n_points, int n_frames, int max_points,
                                   const double *K, const double *M1, const double *M2,
                                   int32_t *counts, int32_t *best, uint8_t *best_mask, int32_t *work,
                                   slamfe_stream_t stream)
{
    if (H < 0 || n_frames < 0 || max_points < 0 || n_points < 0) return SLAMFE_EINVAL;
    if (n_frames == 0) return 0;
    if (!best || !work || !K || !M1 || !M2) return SLAMFE_EINVAL;
    if (H > 0 && (!T || !counts)) return SLAMFE_EINVAL;
    if (!pt_off && n_frames != 1) return SLAMFE_EINVAL;
    if (pt_cnt && !pt_off) return SLAMFE_EINVAL;
    if (!pt_off) max_points = n_points;
    if (max_points > 0 && (!pts || !l_pix || !r_pix || !best_mask)) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SLAMFE_CUDA_OK(cudaMemsetAsync(work, 0, sizeof(int32_t) * n_frames, s));
    if (H > 0) SLAMFE_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * static_cast<size_t>(n_frames) * H, s));
    RansacParams p{};
    p.T = T; p.hyp_valid = hyp_valid; p.H = H;
    p.pts = pts; p.l_pix = l_pix; p.r_pix = r_pix; p.pt_off = pt_off; p.pt_cnt = pt_cnt; p.n_points = n_points;
    for (int k = 0; k < 9; ++k) p.cam.K[k] = K[k];
    for (int k = 0; k < 12; ++k) { p.cam.M1[k] = M1[k]; p.cam.M2[k] = M2[k]; }
    p.counts = counts; p.best = best; p.work = work; p.best_mask = best_mask;
    // hypothesis splits: as few as still give every SM ~8 CTAs' worth of work in flight and to come
    const int tiles = max(1, (max_points + RS_TILE - 1) / RS_TILE), n_chunks = max(1, (H + RS_HC - 1) / RS_HC);
    const long long want = 16LL * sm_count();
    const long long base = static_cast<long long>(tiles) * n_frames;
    long long splits_ll = (want + base - 1) / base;
    if (splits_ll < 1) splits_ll = 1;
    if (splits_ll > n_chunks) splits_ll = n_chunks;
    const int splits = static_cast<int>(splits_ll);
    const dim3 grid(tiles, splits, n_frames);
    if (grid.z > 65535u) return SLAMFE_ERANGE;
    ransac_score_kernel<<<grid, RS_THREADS, 0, s>>>(p);
    return launch_status();
}
