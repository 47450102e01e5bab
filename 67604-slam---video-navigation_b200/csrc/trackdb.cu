// Track ids of a whole sequence on the device — the bookkeeping half of TrackingDB.add_frame
// (final_project/backend/database/tracking_database.py:273-337) as data-parallel passes.
//
// add_frame walks the inlier forward matches (previous feature i -> current feature j) of one frame
// pair in ascending i: a previous feature without a track starts a new one (issue_trackId, :196-198,
// ids are consecutive in processing order), the current feature inherits the previous feature's track
// (:324-329).  create_db only flags MUTUAL matches as inliers (database.py:67-85), so the matches of a
// pair are one-to-one and the "better match to the same feature" branch (:309-322) never fires: tracks
// are simple chains.  Therefore
//   * head of a track = a link with an inlier successor and no inlier predecessor;
//   * its id = number of heads in earlier frames + its rank among the heads of its frame (ascending
//     feature index) — a prefix sum, exactly the order issue_trackId hands ids out;
//   * every other link of the chain walks its predecessors back to the head and copies the id.
// Four small launches over the link rows of the sequence (HBM-trivial: ~20 B per row).
#include "common.cuh"

namespace slamfe {
namespace {

constexpr int TI_THREADS = 256;

struct TrackIdParams {
    const uint32_t *fwd_keys;   // (L,) rows of frame f (link order): compact best key -> link index in frame f + 1
    const uint8_t *inlier_fwd;  // (L,) in_prev_cur of database.py:84-85
    const int32_t *l_off, *n_links;
    int n_frames;
    int32_t *pred, *rank, *head_cnt, *head_base, *track_id, *n_tracks;
};

__device__ __forceinline__ bool has_successor(const TrackIdParams &p, int f, int row, int &j)
{
    if (f >= p.n_frames - 1 || !p.inlier_fwd[row]) return false;
    const uint32_t k = p.fwd_keys[row];
    if (k == KEY_NONE) return false;
    j = static_cast<int>(k & KEY_IDX_MASK);
    return j < p.n_links[f + 1];
}

__global__ void __launch_bounds__(TI_THREADS) track_pred_kernel(const TrackIdParams p)
{
    const int f = blockIdx.x;
    const int l0 = p.l_off[f], n = p.n_links[f], l1 = p.l_off[f + 1];
    for (int i = threadIdx.x; i < n; i += TI_THREADS) {
        int j;
        if (has_successor(p, f, l0 + i, j)) p.pred[l1 + j] = i;
    }
}

__global__ void __launch_bounds__(TI_THREADS) track_heads_kernel(const TrackIdParams p)
{
    __shared__ int s_warp[TI_THREADS / 32];
    __shared__ int s_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l0 = p.l_off[f], n = p.n_links[f], cap = p.l_off[f + 1] - l0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < cap; i0 += TI_THREADS) {
        const int i = i0 + tid;
        int j;
        const bool head = i < n && p.pred[l0 + i] < 0 && has_successor(p, f, l0 + i, j);
        const unsigned b = __ballot_sync(0xFFFFFFFFu, head);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (i < cap) p.rank[l0 + i] = head ? before + __popc(b & ((1u << lane) - 1u)) : -1;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < TI_THREADS / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) p.head_cnt[f] = s_base;
}

__global__ void __launch_bounds__(1024) track_scan_kernel(const TrackIdParams p)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int f0 = 0; f0 < p.n_frames; f0 += 1024) {
        const int f = f0 + tid;
        const int v = f < p.n_frames ? p.head_cnt[f] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (f < p.n_frames) p.head_base[f] = before + x - v;
        __syncthreads();
        if (tid == 1023) s_carry = before + x;
        __syncthreads();
    }
    if (tid == 0) {
        p.head_base[p.n_frames] = s_carry;
        *p.n_tracks = s_carry;
    }
}

__global__ void __launch_bounds__(TI_THREADS) track_assign_kernel(const TrackIdParams p)
{
    const int f = blockIdx.x;
    const int l0 = p.l_off[f], n = p.n_links[f], cap = p.l_off[f + 1] - l0;
    for (int i = threadIdx.x; i < cap; i += TI_THREADS) {
        int id = -1;
        if (i < n) {
            int ff = f, ii = i;
            for (;;) {  // walk the chain back to its head (tracks are a handful of frames long)
                const int pr = p.pred[p.l_off[ff] + ii];
                if (pr < 0) break;
                ii = pr;
                --ff;
            }
            const int r = p.rank[p.l_off[ff] + ii];
            if (r >= 0) id = p.head_base[ff] + r;
        }
        p.track_id[l0 + i] = id;
    }
}

// ---- dense store: the link columns of all frames, frame after frame, without the per-frame padding ----
struct PackDbParams {
    const int32_t *l_off, *r_off, *n_links, *link_src, *match_t, *track_id;
    const float2 *pl, *pr;
    const uint8_t *feat;   // (L, 64) zero-padded descriptors of the links (stereo_links_kernel)
    int desc_bytes;
    int32_t *link_off;     // (F + 1,) exclusive prefix sum of n_links
    float *x_left, *x_right;
    double *y;
    uint8_t *feat_out;     // (N, desc_bytes)
    int32_t *track_out;
};

__global__ void __launch_bounds__(1024) link_offsets_kernel(const int32_t *n_links, int n_frames, int32_t *link_off)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int f0 = 0; f0 < n_frames; f0 += 1024) {
        const int f = f0 + tid;
        const int v = f < n_frames ? n_links[f] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (f < n_frames) link_off[f] = before + x - v;
        __syncthreads();
        if (tid == 1023) s_carry = before + x;
        __syncthreads();
    }
    if (tid == 0) link_off[n_frames] = s_carry;
}

__global__ void __launch_bounds__(TI_THREADS) pack_db_kernel(const PackDbParams p)
{
    const int f = blockIdx.x;
    const int l0 = p.l_off[f], r0 = p.r_off[f], n = p.n_links[f];
    const size_t o0 = static_cast<size_t>(p.link_off[f]);
    for (int k = threadIdx.x; k < n; k += TI_THREADS) {
        const int src = p.link_src[l0 + k];
        const float2 a = p.pl[l0 + src];
        const float2 b = p.pr[r0 + p.match_t[l0 + src]];
        p.x_left[o0 + k] = a.x;
        p.x_right[o0 + k] = b.x;
        p.y[o0 + k] = (static_cast<double>(a.y) + static_cast<double>(b.y)) / 2.0;  // tracking_database.py:243
        p.track_out[o0 + k] = p.track_id ? p.track_id[l0 + k] : -1;
    }
    // descriptors: the frame's n * desc_bytes output bytes are contiguous; the source rows are 64 bytes apart
    const int db = p.desc_bytes;
    const size_t total = static_cast<size_t>(n) * db;
    uint8_t *dst = p.feat_out + o0 * db;
    const uint8_t *srcb = p.feat + static_cast<size_t>(l0) * 64;
    for (size_t b = threadIdx.x; b < total; b += TI_THREADS) {
        const int row = static_cast<int>(b / db), col = static_cast<int>(b - static_cast<size_t>(row) * db);
        dst[b] = srcb[static_cast<size_t>(row) * 64 + col];
    }
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_pack_db(const int32_t *l_off, const int32_t *r_off, const int32_t *n_links, const int32_t *link_src,
                              const int32_t *match_t, const float *pts_left, const float *pts_right, const uint8_t *feat,
                              int desc_bytes, const int32_t *track_id, int n_frames, int32_t *link_off, float *x_left,
                              float *x_right, double *y, uint8_t *feat_out, int32_t *track_out, slamfe_stream_t stream)
{
    if (n_frames < 0 || desc_bytes < 1 || desc_bytes > 64) return SLAMFE_EINVAL;
    if (!link_off) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_frames == 0) {
        SLAMFE_CUDA_OK(cudaMemsetAsync(link_off, 0, sizeof(int32_t), s));
        return 0;
    }
    if (!l_off || !r_off || !n_links || !link_src || !match_t || !pts_left || !pts_right || !feat || !x_left ||
        !x_right || !y || !feat_out || !track_out)
        return SLAMFE_EINVAL;
    PackDbParams p{};
    p.l_off = l_off; p.r_off = r_off; p.n_links = n_links; p.link_src = link_src; p.match_t = match_t;
    p.track_id = track_id;
    p.pl = reinterpret_cast<const float2 *>(pts_left);
    p.pr = reinterpret_cast<const float2 *>(pts_right);
    p.feat = feat; p.desc_bytes = desc_bytes; p.link_off = link_off;
    p.x_left = x_left; p.x_right = x_right; p.y = y; p.feat_out = feat_out; p.track_out = track_out;
    link_offsets_kernel<<<1, 1024, 0, s>>>(n_links, n_frames, link_off);
    pack_db_kernel<<<n_frames, TI_THREADS, 0, s>>>(p);
    return launch_status();
}

extern "C" int slamfe_track_ids(const uint32_t *fwd_keys, const uint8_t *inlier_fwd, const int32_t *l_off,
                                const int32_t *n_links, int n_frames, int64_t rows_total, int32_t *pred, int32_t *rank,
                                int32_t *head_cnt, int32_t *head_base, int32_t *track_id, int32_t *n_tracks,
                                slamfe_stream_t stream)
{
    if (n_frames < 0 || rows_total < 0) return SLAMFE_EINVAL;
    if (!n_tracks) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_frames == 0 || rows_total == 0) {
        SLAMFE_CUDA_OK(cudaMemsetAsync(n_tracks, 0, sizeof(int32_t), s));
        if (head_base && n_frames >= 0) SLAMFE_CUDA_OK(cudaMemsetAsync(head_base, 0, sizeof(int32_t) * (n_frames + 1), s));
        return 0;
    }
    if (!fwd_keys || !inlier_fwd || !l_off || !n_links || !pred || !rank || !head_cnt || !head_base || !track_id)
        return SLAMFE_EINVAL;
    if (n_frames > 65535 * 32) return SLAMFE_ERANGE;
    TrackIdParams p{};
    p.fwd_keys = fwd_keys;
    p.inlier_fwd = inlier_fwd; p.l_off = l_off; p.n_links = n_links; p.n_frames = n_frames;
    p.pred = pred; p.rank = rank; p.head_cnt = head_cnt; p.head_base = head_base; p.track_id = track_id;
    p.n_tracks = n_tracks;
    SLAMFE_CUDA_OK(cudaMemsetAsync(pred, 0xFF, sizeof(int32_t) * static_cast<size_t>(rows_total), s));
    if (n_frames > 1) track_pred_kernel<<<n_frames - 1, TI_THREADS, 0, s>>>(p);
    track_heads_kernel<<<n_frames, TI_THREADS, 0, s>>>(p);
    track_scan_kernel<<<1, 1024, 0, s>>>(p);
    track_assign_kernel<<<n_frames, TI_THREADS, 0, s>>>(p);
    return launch_status();
}
