// Rectified-stereo row filter and link construction — sm_100a.
//
// Replaces extract_inliers_outliers (final_project/algorithms/matching.py:48-69) and, in the
// batched form, the whole per-frame epilogue of database.py:12-27: crossCheck (matching.py:44),
// row filter, TrackingDB.create_links (backend/database/tracking_database.py:224-246) and the
// features[is_valid] compaction.  HBM-bound byte/float shuffling: one CTA per frame, ballot-based
// ordered compaction (links come out ascending in left keypoint index, which is exactly the
// order cv2's crossCheck list + the reference's loops produce).
#include "common.cuh"

namespace slamfe {
namespace {

// matching.py:62-63 on float32 keypoints, evaluated in double like the reference's Python floats.
__device__ __forceinline__ bool stereo_inlier(float xl, float yl, float xr, float yr)
{
    const double dyl = static_cast<double>(yl), dyr = static_cast<double>(yr);
    const double dxl = static_cast<double>(xl), dxr = static_cast<double>(xr);
    return (fabs(dyl - dyr) < 2.0) && (dxl > dxr + 2.0);
}

__global__ void stereo_filter_kernel(const float2 *__restrict__ pl, const float2 *__restrict__ pr,
                                     const int32_t *__restrict__ mq, const int32_t *__restrict__ mt, int n,
                                     uint8_t *__restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = pl[mq[i]], b = pr[mt[i]];
    mask[i] = stereo_inlier(a.x, a.y, b.x, b.y) ? 1 : 0;
}

constexpr int SL_THREADS = 512;
constexpr int SL_TILES = 16;                          // tiles of SL_THREADS rows per pass
constexpr int SL_PASS = SL_TILES * SL_THREADS;        // 8192 rows: one pass covers any frame of the workload
constexpr int SL_WARPS = SL_THREADS / 32;
static_assert(SL_TILES * SL_WARPS <= SL_THREADS, "one scan element per thread");

struct StereoLinksParams {
    const uint2 *row_keys;
    const uint32_t *col_keys;
    const int32_t *l_off, *r_off, *l_cnt, *r_cnt;
    const float2 *pl, *pr;
    const uint8_t *desc;
    int l_stride, desc_bytes;
    int32_t *match_t, *n_matches, *n_links, *link_src;
    float *links;
    uint8_t *feat;
};

// One CTA per frame, three phases per pass of up to 8192 left rows, five barriers per pass (the first
// version walked 512-row tiles with four barriers each and was bound by their latency):
//   1. every thread classifies its rows of all tiles (crossCheck: row key -> column key -> mutual; row filter on
//      the two keypoints) — 16 independent dependent-load chains per thread in flight — writes match_t and
//      leaves one ballot count per (tile, warp) in shared memory;
//   2. after an exclusive scan of the 256 counts, the good rows write link_src / links at their rank (ascending
//      left keypoint index: the order cv2's crossCheck list + the reference's loops produce) and their row
//      number into a shared list;
//   3. 16 threads per link copy the descriptors of the listed rows into zero-padded 64-byte rows.
__global__ void __launch_bounds__(SL_THREADS) stereo_links_kernel(const StereoLinksParams p)
{
    __shared__ int s_cnt[SL_TILES * SL_WARPS];    // good rows per (tile, warp), then their exclusive prefix
    __shared__ int s_wsum[SL_WARPS];
    __shared__ int s_mut[SL_WARPS];
    __shared__ uint16_t s_src[SL_PASS];            // left row (relative to the pass) of the pass' links, in order
    __shared__ int s_total;

    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l0 = p.l_off[f], nl = p.l_cnt ? p.l_cnt[f] : p.l_off[f + 1] - l0;
    const int r0 = p.r_off[f], nr = p.r_cnt ? p.r_cnt[f] : p.r_off[f + 1] - r0;
    int base = 0, n_mut = 0;   // links written so far; mutual matches seen by this warp (lane 0)

    for (int pass0 = 0; pass0 < nl; pass0 += SL_PASS) {
        const int n_tiles = min(SL_TILES, (nl - pass0 + SL_THREADS - 1) / SL_THREADS);
        // ---- 1. classify ----
        uint32_t good_bits = 0;
#pragma unroll 4
        for (int t = 0; t < n_tiles; ++t) {
            const int i = pass0 + t * SL_THREADS + tid;
            bool mutual = false, good = false;
            if (i < nl) {
                int j = -1;
                const uint32_t k = p.row_keys[l0 + i].x;
                if (k != KEY_NONE) {
                    const uint32_t jj = k & KEY_IDX_MASK;
                    if (jj < static_cast<uint32_t>(nr)) {
                        const uint32_t c = p.col_keys[r0 + jj];
                        mutual = (c != KEY_NONE) && ((c & KEY_IDX_MASK) == static_cast<uint32_t>(i));
                        if (mutual) j = static_cast<int>(jj);
                    }
                }
                p.match_t[l0 + i] = j;
                if (mutual) {
                    const float2 a = p.pl[l0 + i], b = p.pr[r0 + j];
                    good = stereo_inlier(a.x, a.y, b.x, b.y);
                }
            }
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, good);
            const uint32_t balm = __ballot_sync(0xFFFFFFFFu, mutual);
            if (good) good_bits |= 1u << t;
            if (lane == 0) {
                s_cnt[t * SL_WARPS + warp] = __popc(bal);
                n_mut += __popc(balm);
            }
        }
        __syncthreads();
        // ---- exclusive scan of the n_tiles * SL_WARPS counts (tile-major: ascending row order) ----
        {
            const int n_el = n_tiles * SL_WARPS;
            const int v = tid < n_el ? s_cnt[tid] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            int woff = 0;
#pragma unroll
            for (int w = 0; w < SL_WARPS; ++w)
                if (w < warp) woff += s_wsum[w];
            if (tid < n_el) s_cnt[tid] = woff + inc - v;
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < SL_WARPS; ++w) tot += s_wsum[w];
                s_total = tot;
            }
            __syncthreads();
        }
        const int total = s_total;
        // ---- 2. links of the good rows, at their rank ----
#pragma unroll 4
        for (int t = 0; t < n_tiles; ++t) {
            const bool good = (good_bits >> t) & 1u;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, good);
            if (good) {
                const int i = pass0 + t * SL_THREADS + tid;
                const int pos = s_cnt[t * SL_WARPS + warp] + __popc(bal & ((1u << lane) - 1u));
                const int j = p.match_t[l0 + i];   // this thread's own store of phase 1
                const float2 a = p.pl[l0 + i], b = p.pr[r0 + j];
                const size_t o = static_cast<size_t>(l0 + base + pos);
                p.link_src[o] = i;
                p.links[3 * o + 0] = a.x;
                p.links[3 * o + 1] = b.x;
                // tracking_database.py:243: (yl + yr) / 2 in double, stored as float
                p.links[3 * o + 2] = static_cast<float>((static_cast<double>(a.y) + static_cast<double>(b.y)) / 2.0);
                s_src[pos] = static_cast<uint16_t>(i - pass0);
            }
        }
        __syncthreads();
        // ---- 3. features[is_valid] ----
        if (p.feat) {
            // 16 threads per link copy one descriptor into a zero-padded 64-byte row; SL_BATCH items per thread
            // are loaded before any is stored, so their global loads (L2 / HBM latency) overlap
            constexpr int SL_BATCH = 8;
            uint32_t *feat32 = reinterpret_cast<uint32_t *>(p.feat) + static_cast<size_t>(l0 + base) * 16;
            for (int e0 = tid; e0 < total * 16; e0 += SL_THREADS * SL_BATCH) {
                uint32_t lo[SL_BATCH], hi[SL_BATCH];
#pragma unroll
                for (int u = 0; u < SL_BATCH; ++u) {
                    const int e = e0 + u * SL_THREADS;
                    lo[u] = hi[u] = 0u;
                    if (e < total * 16) {
                        const int k = e >> 4, w = e & 15;
                        // bytes [4w, 4w + 4) of a row at any alignment: the one or two aligned words that hold
                        // them (a word is only read when it contains a valid descriptor byte)
                        const int valid = min(4, p.desc_bytes - 4 * w);
                        if (valid > 0) {
                            const uintptr_t a = reinterpret_cast<uintptr_t>(
                                p.desc + static_cast<size_t>(l0 + pass0 + s_src[k]) * p.l_stride + 4 * w);
                            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
                            lo[u] = __ldg(s32);
                            if (static_cast<int>(a & 3) + valid > 4) hi[u] = __ldg(s32 + 1);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < SL_BATCH; ++u) {
                    const int e = e0 + u * SL_THREADS;
                    if (e < total * 16) {
                        const int k = e >> 4, w = e & 15;
                        const int valid = min(4, p.desc_bytes - 4 * w);
                        uint32_t v = 0;
                        if (valid > 0) {   // funnel shift by the row's misalignment, tail mask
                            const int mis = static_cast<int>(
                                (reinterpret_cast<uintptr_t>(p.desc) + static_cast<size_t>(l0 + pass0 + s_src[k]) * p.l_stride) & 3);
                            v = __funnelshift_r(lo[u], hi[u], mis * 8);
                            if (valid < 4) v &= (1u << (8 * valid)) - 1u;
                        }
                        feat32[static_cast<size_t>(k) * 16 + w] = v;
                    }
                }
            }
        }
        base += total;
        __syncthreads();   // s_cnt / s_src / s_total are rewritten by the next pass
    }
    // rows of this frame's capacity that hold no link: deterministic filler
    for (int k = base + tid; k < nl; k += SL_THREADS) {
        const size_t o = static_cast<size_t>(l0 + k);
        p.link_src[o] = -1;
        p.links[3 * o + 0] = 0.f;
        p.links[3 * o + 1] = 0.f;
        p.links[3 * o + 2] = 0.f;
    }
    if (lane == 0) s_mut[warp] = n_mut;
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int w = 0; w < SL_WARPS; ++w) m += s_mut[w];
        p.n_links[f] = base;
        p.n_matches[f] = m;
    }
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_stereo_filter(const float *pts_left, const float *pts_right, const int32_t *match_q,
                                    const int32_t *match_t, int n_matches, uint8_t *mask, slamfe_stream_t stream)
{
    if (n_matches < 0) return SLAMFE_EINVAL;
    if (n_matches == 0) return 0;
    if (!pts_left || !pts_right || !match_q || !match_t || !mask) return SLAMFE_EINVAL;
    stereo_filter_kernel<<<(n_matches + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(pts_left), reinterpret_cast<const float2 *>(pts_right), match_q, match_t,
        n_matches, mask);
    return launch_status();
}

extern "C" int slamfe_stereo_links_batched(const uint32_t *row_keys, const uint32_t *col_keys, const int32_t *l_off,
                                           const int32_t *l_cnt, const int32_t *r_off, const int32_t *r_cnt,
                                           int n_frames, const float *pts_left,
                                           const float *pts_right, const uint8_t *desc_left, int l_stride,
                                           int desc_bytes, int32_t *match_t, int32_t *n_matches, int32_t *n_links,
                                           int32_t *link_src, float *links, uint8_t *feat, slamfe_stream_t stream)
{
    if (n_frames < 0) return SLAMFE_EINVAL;
    if (n_frames == 0) return 0;
    if (!row_keys || !col_keys || !l_off || !r_off || !pts_left || !pts_right || !match_t || !n_matches || !n_links ||
        !link_src || !links)
        return SLAMFE_EINVAL;
    if (feat && (!desc_left || desc_bytes <= 0 || desc_bytes > SLAMFE_MAX_DESC_BYTES || l_stride < desc_bytes))
        return SLAMFE_EINVAL;
    StereoLinksParams p{};
    p.row_keys = reinterpret_cast<const uint2 *>(row_keys);
    p.col_keys = col_keys;
    p.l_off = l_off; p.r_off = r_off; p.l_cnt = l_cnt; p.r_cnt = r_cnt;
    p.pl = reinterpret_cast<const float2 *>(pts_left);
    p.pr = reinterpret_cast<const float2 *>(pts_right);
    p.desc = desc_left; p.l_stride = l_stride; p.desc_bytes = desc_bytes;
    p.match_t = match_t; p.n_matches = n_matches; p.n_links = n_links; p.link_src = link_src;
    p.links = links; p.feat = feat;
    stereo_links_kernel<<<n_frames, SL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return launch_status();
}
