// Frame-to-frame tracking glue of the reference's create_db loop, on the device — sm_100a.
//
// Replaces, for ALL consecutive frame pairs of a sequence in one launch (one CTA per pair):
//   * the forward/backward mutual-consistency loop (backend/database/database.py:67-77):
//     forward match j of frame t is kept iff backward[forward[j].trainIdx].trainIdx == j;
//   * the link gather of ransac_pnp_for_tracking_db (algorithms/ransac.py:76-81, :85-88):
//     previous-frame links by queryIdx, current-frame links by trainIdx, pixel arrays
//     (x_left, y) / (x_right, y);
//   * triangulate_links on the previous links (ransac.py:83 -> triangulation.py:41-50), in fp64
//     from the exact Link values: x = float32 keypoint coordinates, y = (yl + yr) / 2 in double
//     (tracking_database.py:243);
//   * calc_ransac_iteration (ransac.py:59-67) from the CURRENT frame's stereo inlier percentage
//     (database.py:26, :80).
// The outputs are the inputs of slamfe_ransac_hypotheses / slamfe_ransac_score, so the whole loop
// body of create_db (database.py:48-87) minus the TrackingDB insertion runs without leaving HBM.
// HBM-bound gather/compaction: ~100 B per mutual match.
#include "common.cuh"
#include "triangulate.cuh"

namespace slamfe {
namespace {

constexpr int TG_THREADS = 256;

struct GatherParams {
    const uint32_t *fwd_keys;   // (L,): rows of frame t (link order), compact best key -> link index in frame t+1
    const uint32_t *bwd_keys;   // (L,):   rows of frame t+1, best key -> link index in frame t
    const int32_t *l_off, *r_off, *n_links, *n_matches;
    const float2 *pl, *pr;      // keypoints (x, y) of the left / right images
    const int32_t *link_src;    // (L,): left keypoint index of the k-th link of a frame
    const int32_t *match_t;     // (L,): mutual right keypoint index per left keypoint
    Cams cam;
    int h_max;
    int32_t *good_j, *good_t, *n_good, *n_hyp, *n_hyp_full;
    double *pts, *lpix, *rpix;
};

// Link k of frame f as the reference's Link(x_left, x_right, y): exact double values.
__device__ __forceinline__ void link_values(const GatherParams &p, int f, int k, double &xl, double &xr, double &y)
{
    const int l0 = p.l_off[f], r0 = p.r_off[f];
    const int src = p.link_src[l0 + k];
    const float2 a = p.pl[l0 + src];
    const float2 b = p.pr[r0 + p.match_t[l0 + src]];
    xl = static_cast<double>(a.x);
    xr = static_cast<double>(b.x);
    y = (static_cast<double>(a.y) + static_cast<double>(b.y)) / 2.0;  // tracking_database.py:243
}

// ransac.py:59-67 (SUCCESS_PROBABILITY = 0.9999999999, minimal set of 4)
__device__ int ransac_iterations(double inliers_percent)
{
    const double suc_prob = 0.9999999999;
    const double outliers_prob = 1.0 - (inliers_percent / 100.0) + 0.0000000001;
    const double w = 1.0 - outliers_prob;
    const double it = log(1.0 - suc_prob) / log(1.0 - w * w * w * w);
    if (!(it == it) || it > 2.0e9) return 0x7FFFFFFF;
    return static_cast<int>(it) + 1;
}

__global__ void __launch_bounds__(TG_THREADS) track_gather_kernel(const GatherParams p)
{
    __shared__ int warp_cnt[TG_THREADS / 32];
    __shared__ int s_base;
    const int pair = blockIdx.x;  // frames (pair, pair + 1)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l0 = p.l_off[pair], l1 = p.l_off[pair + 1];
    const int n_prev = p.n_links[pair], n_cur = p.n_links[pair + 1];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_prev; c0 += TG_THREADS) {
        const int j = c0 + tid;
        bool good = false;
        int t = -1;
        if (j < n_prev) {
            const uint32_t k = p.fwd_keys[l0 + j];
            if (k != KEY_NONE) {
                t = static_cast<int>(k & KEY_IDX_MASK);
                if (t < n_cur) {
                    const uint32_t b = p.bwd_keys[l1 + t];
                    good = (b != KEY_NONE) && (static_cast<int>(b & KEY_IDX_MASK) == j);  // database.py:71
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, good);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TG_THREADS / 32; ++w) {
            if (w < warp) woff += warp_cnt[w];
            total += warp_cnt[w];
        }
        const int base = s_base;
        if (good) {
            const size_t o = static_cast<size_t>(l0) + base + woff + __popc(bal & ((1u << lane) - 1u));
            p.good_j[o] = j;
            p.good_t[o] = t;
            double xl, xr, y, X, Y, Z;
            link_values(p, pair, j, xl, xr, y);
            triangulate_link(p.cam, xl, xr, y, X, Y, Z);  // ransac.py:83
            p.pts[3 * o] = X; p.pts[3 * o + 1] = Y; p.pts[3 * o + 2] = Z;
            link_values(p, pair + 1, t, xl, xr, y);
            p.lpix[2 * o] = xl; p.lpix[2 * o + 1] = y;    // ransac.py:85-88
            p.rpix[2 * o] = xr; p.rpix[2 * o + 1] = y;
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0) {
        p.n_good[pair] = s_base;
        const int nm = p.n_matches[pair + 1];
        int it = 0;
        if (nm > 0) it = ransac_iterations(100.0 * (static_cast<double>(n_cur) / static_cast<double>(nm)));
        p.n_hyp[pair] = min(it, p.h_max);
        if (p.n_hyp_full) p.n_hyp_full[pair] = it;  // it > h_max: this pair's RANSAC is truncated
    }
}

__global__ void scatter_inliers_kernel(const uint8_t *__restrict__ best_mask, const int32_t *__restrict__ good_j,
                                       const int32_t *__restrict__ l_off, const int32_t *__restrict__ n_good,
                                       const int32_t *__restrict__ best, uint8_t *__restrict__ inlier_fwd)
{
    const int pair = blockIdx.x;
    const int l0 = l_off[pair], n = n_good[pair];
    // ransac.py:92,113 + database.py:82: no hypothesis with > 0 inliers -> best_matches_idx is None and
    // good_idx[None] selects EVERY mutual match
    const bool none = best[2 * pair] < 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        if (none || best_mask[l0 + k]) inlier_fwd[l0 + good_j[l0 + k]] = 1;
}

// Loop-closure candidate form of the gather (backend/loop/loop_closure.py:405-436 -> ransac.py:132-146):
// candidate p matches ALL links of keyframe A (queries, in order) against keyframe B; the 3-D points
// are keyframe A's triangulated links, the pixel arrays keyframe B's links at trainIdx.
__global__ void pairs_gather_kernel(const uint32_t *__restrict__ keys, const int32_t *__restrict__ q_off,
                                    const int32_t *__restrict__ q_cnt, const int32_t *__restrict__ t_off,
                                    const int32_t *__restrict__ out_off, const double *__restrict__ kf_pts,
                                    const double *__restrict__ kf_links, double *__restrict__ pts,
                                    double *__restrict__ lpix, double *__restrict__ rpix)
{
    const int p = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q_cnt[p]) return;
    const size_t o = static_cast<size_t>(out_off[p]) + j;
    const size_t a = static_cast<size_t>(q_off[p]) + j;
    pts[3 * o] = kf_pts[3 * a];
    pts[3 * o + 1] = kf_pts[3 * a + 1];
    pts[3 * o + 2] = kf_pts[3 * a + 2];
    const uint32_t k = keys[o];
    double xl, xr, y;
    if (k == KEY_NONE) {  // empty train keyframe: a correspondence that can never agree
        xl = xr = y = __longlong_as_double(0x7FF8000000000000ll);
    } else {
        const size_t b = static_cast<size_t>(t_off[p]) + (k & KEY_IDX_MASK);
        xl = kf_links[3 * b];
        xr = kf_links[3 * b + 1];
        y = kf_links[3 * b + 2];
    }
    lpix[2 * o] = xl; lpix[2 * o + 1] = y;
    rpix[2 * o] = xr; rpix[2 * o + 1] = y;
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_pairs_gather(const uint32_t *keys, const int32_t *q_off, const int32_t *q_cnt,
                                   const int32_t *t_off, const int32_t *out_off, int n_problems, int max_nq,
                                   const double *kf_pts, const double *kf_links, double *pts, double *lpix,
                                   double *rpix, slamfe_stream_t stream)
{
    if (n_problems < 0 || max_nq < 0) return SLAMFE_EINVAL;
    if (n_problems == 0 || max_nq == 0) return 0;
    if (!keys || !q_off || !q_cnt || !t_off || !out_off || !kf_pts || !kf_links || !pts || !lpix || !rpix)
        return SLAMFE_EINVAL;
    if (n_problems > 65535) return SLAMFE_ERANGE;
    const dim3 grid((max_nq + 255) / 256, n_problems);
    pairs_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, q_off, q_cnt, t_off, out_off, kf_pts,
                                                                             kf_links, pts, lpix, rpix);
    return launch_status();
}

extern "C" int slamfe_track_gather(const uint32_t *fwd_keys, const uint32_t *bwd_keys, const int32_t *l_off,
                                   const int32_t *r_off, const int32_t *n_links, const int32_t *n_matches,
                                   const float *pts_left, const float *pts_right, const int32_t *link_src,
                                   const int32_t *match_t, int n_pairs, const double *P, const double *Q, int h_max,
                                   int32_t *good_j, int32_t *good_t, int32_t *n_good, int32_t *n_hyp,
                                   int32_t *n_hyp_full, double *pts, double *lpix, double *rpix,
                                   slamfe_stream_t stream)
{
    if (n_pairs < 0 || h_max < 0) return SLAMFE_EINVAL;
    if (n_pairs == 0) return 0;
    if (!fwd_keys || !bwd_keys || !l_off || !r_off || !n_links || !n_matches || !pts_left || !pts_right ||
        !link_src || !match_t || !good_j || !good_t || !n_good || !n_hyp || !pts || !lpix || !rpix)
        return SLAMFE_EINVAL;
    GatherParams p{};
    const int rc = load_cams(P, Q, p.cam);
    if (rc) return rc;
    for (int k = 4; k < 12; ++k)
        if (p.cam.P[k] != p.cam.Q[k]) return SLAMFE_EINVAL;  // links need a shared-row stereo pair
    p.fwd_keys = fwd_keys;
    p.bwd_keys = bwd_keys;
    p.l_off = l_off; p.r_off = r_off; p.n_links = n_links; p.n_matches = n_matches;
    p.pl = reinterpret_cast<const float2 *>(pts_left);
    p.pr = reinterpret_cast<const float2 *>(pts_right);
    p.link_src = link_src; p.match_t = match_t; p.h_max = h_max;
    p.good_j = good_j; p.good_t = good_t; p.n_good = n_good; p.n_hyp = n_hyp; p.n_hyp_full = n_hyp_full;
    p.pts = pts; p.lpix = lpix; p.rpix = rpix;
    track_gather_kernel<<<n_pairs, TG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return launch_status();
}

extern "C" int slamfe_scatter_inliers(const uint8_t *best_mask, const int32_t *good_j, const int32_t *l_off,
                                      const int32_t *n_good, const int32_t *best, int n_pairs, uint8_t *inlier_fwd,
                                      int64_t rows_total, slamfe_stream_t stream)
{
    if (n_pairs < 0 || rows_total < 0) return SLAMFE_EINVAL;
    if (rows_total == 0) return 0;
    if (!inlier_fwd) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SLAMFE_CUDA_OK(cudaMemsetAsync(inlier_fwd, 0, static_cast<size_t>(rows_total), s));
    if (n_pairs == 0) return 0;
    if (!best_mask || !good_j || !l_off || !n_good || !best) return SLAMFE_EINVAL;
    scatter_inliers_kernel<<<n_pairs, 256, 0, s>>>(best_mask, good_j, l_off, n_good, best, inlier_fwd);
    return launch_status();
}
