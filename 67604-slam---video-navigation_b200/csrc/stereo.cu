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

constexpr int SL_THREADS = 512;   // one CTA per frame: the tile loop is latency-bound (4 barriers per tile), so fewer, wider tiles

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

__global__ void __launch_bounds__(SL_THREADS) stereo_links_kernel(const StereoLinksParams p)
{
    __shared__ int warp_cnt[SL_THREADS / 32];
    __shared__ int warp_mut[SL_THREADS / 32];
    __shared__ int chunk_src[SL_THREADS];  // left rows of this chunk's links, in order
    __shared__ int s_base, s_mut;

    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l0 = p.l_off[f], nl = p.l_cnt ? p.l_cnt[f] : p.l_off[f + 1] - l0;
    const int r0 = p.r_off[f], nr = p.r_cnt ? p.r_cnt[f] : p.r_off[f + 1] - r0;
    if (tid == 0) { s_base = 0; s_mut = 0; }
    __syncthreads();

    for (int c0 = 0; c0 < nl; c0 += SL_THREADS) {
        const int i = c0 + tid;
        bool mutual = false, good = false;
        int j = -1;
        float xl = 0.f, xr = 0.f, y = 0.f;
        if (i < nl) {
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
                xl = a.x;
                xr = b.x;
                // tracking_database.py:243: (yl + yr) / 2 in double, stored as float
                y = static_cast<float>((static_cast<double>(a.y) + static_cast<double>(b.y)) / 2.0);
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, good);
        const uint32_t balm = __ballot_sync(0xFFFFFFFFu, mutual);
        if (lane == 0) {
            warp_cnt[warp] = __popc(bal);
            warp_mut[warp] = __popc(balm);
        }
        __syncthreads();
        int woff = 0, total = 0, mtotal = 0;
#pragma unroll
        for (int w = 0; w < SL_THREADS / 32; ++w) {
            if (w < warp) woff += warp_cnt[w];
            total += warp_cnt[w];
            mtotal += warp_mut[w];
        }
        const int base = s_base;
        if (good) {
            const int pos = woff + __popc(bal & ((1u << lane) - 1u));
            const size_t o = static_cast<size_t>(l0 + base + pos);
            p.link_src[o] = i;
            p.links[3 * o + 0] = xl;
            p.links[3 * o + 1] = xr;
            p.links[3 * o + 2] = y;
            chunk_src[pos] = i;
        }
        __syncthreads();
        if (p.feat) {
            // 16 threads per link copy one descriptor into a zero-padded 64-byte row
            for (int e = tid; e < total * 16; e += SL_THREADS) {
                const int k = e >> 4, w = e & 15;
                const uint8_t *src = p.desc + static_cast<size_t>(l0 + chunk_src[k]) * p.l_stride + 4 * w;
                // bytes [4w, 4w + 4) of a row at any alignment: the one or two aligned words that hold them
                // (a word is only read when it contains a valid descriptor byte), funnel shift, tail mask
                const int valid = min(4, p.desc_bytes - 4 * w);
                uint32_t v = 0;
                if (valid > 0) {
                    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
                    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
                    const int mis = static_cast<int>(a & 3);
                    const uint32_t lo = __ldg(s32);
                    const uint32_t hi = (mis + valid > 4) ? __ldg(s32 + 1) : 0u;
                    v = __funnelshift_r(lo, hi, mis * 8);
                    if (valid < 4) v &= (1u << (8 * valid)) - 1u;
                }
                reinterpret_cast<uint32_t *>(p.feat)[static_cast<size_t>(l0 + base + k) * 16 + w] = v;
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_base = base + total;
            s_mut += mtotal;
        }
        __syncthreads();
    }
    // rows of this frame's capacity that hold no link: deterministic filler
    for (int k = s_base + tid; k < nl; k += SL_THREADS) {
        const size_t o = static_cast<size_t>(l0 + k);
        p.link_src[o] = -1;
        p.links[3 * o + 0] = 0.f;
        p.links[3 * o + 1] = 0.f;
        p.links[3 * o + 2] = 0.f;
    }
    if (tid == 0) {
        p.n_links[f] = s_base;
        p.n_matches[f] = s_mut;
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
