// Brute-force Hamming top-2 matcher for 61-byte AKAZE (MLDB) descriptors — sm_100a.
//
// Replaces cv2.BFMatcher(NORM_HAMMING).match / knnMatch(k=2) / crossCheck at the reference call
// sites final_project/backend/database/database.py:54-55, backend/loop/loop_closure.py:422,
// final_project/algorithms/matching.py:15,44 and VAN_ex/code/ex1.py:189-190.
//
// Design (integer-pipe bound; no tensor cores — this is XOR+POPC, not a dense contraction):
//   * grid = (query tiles, train slices, problems): one launch covers a ragged batch of frame
//     pairs; the hardware CTA scheduler is the work queue.
//   * every thread keeps up to RMAX query descriptors (16 x u32 each) and their running best keys
//     in registers for the whole train sweep;
//   * the train rows stream through shared memory in 128-row stages: a 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier) lands the raw 61-byte rows, the CTA re-aligns them to 64-byte
//     rows (funnel shift) into a double-buffered tile, and all lanes read the same train row
//     with broadcast LDS.128;
//   * prefix-XOR carry-save distance (below): 16 LOP3 + 11 LOP3 + 7 POPC per descriptor pair
//     instead of 16 LOP3 + 16 POPC — POPC runs on the XU pipe at 16 lanes/clk/SM and was the
//     limiter (ncu: XU 92 %), LOP3 runs on the ALU pipe at 64 lanes/clk/SM;
//   * key = (distance << 22) | index, so unsigned min == cv2's first-minimum tie-break; the
//     per-train-row (column) minimum for crossCheck / the backward match comes from the same
//     pass: lane-min, one warp reduction per row, per-warp shared slots, one global atomicMin
//     per CTA and train row.
#include <stdlib.h>

#include "common.cuh"
#include "hamming_core.cuh"
#include "hamming_params.cuh"

namespace slamfe {

namespace {

constexpr int TS = 128;  // train rows per shared-memory stage
constexpr int RAW_BYTES = TS * SLAMFE_MAX_DESC_BYTES + 16;

// One shared-memory stage (rows train rows) against the NR query rows each lane holds.
template <int NR, int RMAX, bool COL, bool TOP2, int CS>
__device__ __forceinline__ void sweep_stage(const uint4 *__restrict__ tl, int rows, uint32_t jbase,
                                            const uint32_t (&q)[RMAX][W], uint32_t (&b1)[RMAX], uint32_t (&b2)[RMAX],
                                            const uint32_t (&colbias)[RMAX], uint32_t colmin_in, int lane)
{
    uint32_t jj = jbase;
    uint32_t colmin_w;  // pinned in a register: ptxas otherwise rematerialises the address every row
    asm volatile("mov.u32 %0, %1;" : "=r"(colmin_w) : "r"(colmin_in));
#pragma unroll 2
    for (int j = 0; j < rows; ++j, ++jj) {
        const uint4 t0 = tl[4 * j + 0], t1 = tl[4 * j + 1], t2 = tl[4 * j + 2], t3 = tl[4 * j + 3];
        uint32_t ck = KEY_NONE;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const uint32_t dk = hamming16_key<CS>(q[r], t0, t1, t2, t3);
            const uint32_t key = dk + jj;
            if (TOP2) {
                const uint32_t hi = max(b1[r], key);
                b2[r] = min(b2[r], hi);
            }
            b1[r] = min(b1[r], key);
            if (COL) ck = min(ck, dk + colbias[r]);
        }
        if (COL) {
            const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, ck);
            // explicit shared-space store: one address register + immediate, no generic-pointer
            // conversion inside the loop (ptxas rematerialised it every row under register pressure)
            if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(colmin_w + 4u * j), "r"(m) : "memory");
        }
    }
}

template <int N, int RMAX, bool COL, bool TOP2, int CS>
__device__ __forceinline__ void sweep_dispatch(int nr_w, const uint4 *__restrict__ tl, int rows, uint32_t jbase,
                                               const uint32_t (&q)[RMAX][W], uint32_t (&b1)[RMAX],
                                               uint32_t (&b2)[RMAX], const uint32_t (&colbias)[RMAX],
                                               uint32_t colmin_w, int lane)
{
    if (nr_w == N)
        sweep_stage<N, RMAX, COL, TOP2, CS>(tl, rows, jbase, q, b1, b2, colbias, colmin_w, lane);
    else if constexpr (N > 1)
        sweep_dispatch<N - 1, RMAX, COL, TOP2, CS>(nr_w, tl, rows, jbase, q, b1, b2, colbias, colmin_w, lane);
}

// grid = (query tiles of RMAX*THREADS rows, train slices, problems).  Query rows are dealt to
// threads in slabs of THREADS rows (row = tile + r*THREADS + tid), so a warp of a partial tile
// only sweeps the slabs that hold rows (nr_w of them) and an all-empty warp only helps staging.
template <int THREADS, int RMAX, bool COL, bool TOP2, int CS>
__global__ void __launch_bounds__(THREADS, (RMAX >= 3 ? 2 : 3)) hamming_top2_kernel(const HammingParams p)
{
    constexpr int TQ = RMAX * THREADS;
    constexpr int NWARP = THREADS / 32;
    __shared__ alignas(128) uint32_t tile[2][TS * W];
    __shared__ alignas(128) uint8_t raw[2][RAW_BYTES];
    __shared__ uint32_t colmin_s[COL ? 2 : 1][COL ? NWARP : 1][COL ? TS : 1];
    __shared__ alignas(8) uint64_t mbar[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int prob = blockIdx.z;

    int q_row0 = 0, nq = p.nq, t_row0 = 0, nt = p.nt;
    if (p.q_off) {
        q_row0 = p.q_off[prob];
        nq = p.q_cnt ? p.q_cnt[prob] : p.q_off[prob + 1] - q_row0;
    }
    if (p.t_off) {
        t_row0 = p.t_off[prob];
        nt = p.t_cnt ? p.t_cnt[prob] : p.t_off[prob + 1] - t_row0;
    }
    const int qt0 = blockIdx.x * TQ;
    const int tb = blockIdx.y * p.t_slice;
    if (qt0 >= nq || tb >= nt) return;  // CTA-uniform; outputs were pre-set to KEY_NONE
    const int te = min(nt, tb + p.t_slice);
    const int n_stage = (te - tb + TS - 1) / TS;
    // warps that hold at least one query row (the others never write column minima)
    const int n_active_warps = min(NWARP, (nq - qt0 + 31) / 32);

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint8_t *t_base = p.t + static_cast<size_t>(t_row0) * p.t_stride;
    auto stage_rows = [&](int s) { return min(TS, te - (tb + s * TS)); };
    auto stage_src = [&](int s) { return t_base + static_cast<size_t>(tb + s * TS) * p.t_stride; };
    auto stage_tma_rows = [&](int s) {
        if (reinterpret_cast<uintptr_t>(stage_src(s)) & 15) return 0;
        const int rows = stage_rows(s);
        return rows - rows % p.tma_quantum;
    };
    auto issue_stage = [&](int s) {  // thread 0 only
        const int trows = stage_tma_rows(s);
        if (trows > 0) {
            const uint32_t bytes = static_cast<uint32_t>(trows) * p.t_stride;
            mbar_arrive_expect_tx(&mbar[s & 1], bytes);
            tma_load_1d(raw[s & 1], stage_src(s), bytes, &mbar[s & 1]);
        }
    };
    auto flush_cols = [&](int s) {  // threads 0..TS-1: column minima of stage s -> global
        if (tid < stage_rows(s)) {
            uint32_t v = KEY_NONE;
            for (int w = 0; w < n_active_warps; ++w) v = min(v, colmin_s[s & 1][w][tid]);
            atomicMin(p.col_keys + t_row0 + tb + s * TS + tid, v);
        }
    };
    if (tid == 0) {
        issue_stage(0);
        if (n_stage > 1) issue_stage(1);
    }

    // ---- query descriptors -> registers, prefix form (held for the whole sweep) ----
    // A lane past the last row of a partial slab mirrors row nq-1: its keys duplicate a real
    // lane's, so the column minima need no masking; only its row result is not written.
    const int warp_row0 = qt0 + (tid & ~31);
    int nr_w = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
        if (warp_row0 + r * THREADS < nq) nr_w = r + 1;
    uint32_t q[RMAX][W];
    uint32_t b1[RMAX], b2[RMAX], colbias[RMAX];
    int qrow[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        qrow[r] = qt0 + tid + r * THREADS;
        b1[r] = KEY_NONE;
        b2[r] = KEY_NONE;
        const int src_row = min(qrow[r], nq - 1);
        colbias[r] = static_cast<uint32_t>(src_row);
        if (r < nr_w) {
            load_desc_global(p.q + static_cast<size_t>(q_row0 + src_row) * p.q_stride, p.desc_bytes, q[r]);
            to_prefix_form(q[r]);
        } else {
#pragma unroll
            for (int k = 0; k < W; ++k) q[r][k] = 0;
        }
    }

    uint32_t phase = 0;  // bit b = parity to wait for on mbar[b]
    for (int s = 0; s < n_stage; ++s) {
        const int b = s & 1;
        const int rows = stage_rows(s);
        const int trows = stage_tma_rows(s);
        if (trows > 0) {
            mbar_wait(&mbar[b], (phase >> b) & 1u);
            phase ^= 1u << b;
        }
        // flush the column minima of stage s-2 (same buffer) before this stage overwrites them
        if (COL && s >= 2) flush_cols(s - 2);
        // ---- re-align raw rows (any stride) to 64-byte rows in prefix form ----
        {
            const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(raw[b]);
            const uint8_t *src = stage_src(s);
            for (int i = tid; i < TS * W; i += THREADS) {  // 16 consecutive lanes = one row
                const int r = i >> 4, k = i & 15;
                uint32_t v = 0;
                if (r < trows) {
                    const int o = r * p.t_stride + 4 * k;
                    v = __funnelshift_r(raw32[o >> 2], raw32[(o >> 2) + 1], (o & 3) * 8) & word_mask(p.desc_bytes, k);
                } else if (r < rows) {
                    const uint8_t *g = src + static_cast<size_t>(r) * p.t_stride + 4 * k;
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb)
                        if (4 * k + bb < p.desc_bytes) v |= static_cast<uint32_t>(__ldg(g + bb)) << (8 * bb);
                }
                uint32_t pre = v;  // inclusive XOR scan over the 16 words of the row
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, pre, o, 16);
                    if (k >= o) pre ^= up;
                }
                tile[b][i] = (k >= 2 && k <= 14 && !(k & 1)) ? pre : v;
            }
        }
        __syncthreads();  // tile[b] complete; raw[b] and tile[b^1] are free again
        if (tid == 0 && s + 2 < n_stage) issue_stage(s + 2);

        // ---- sweep: every lane reads the same train row (broadcast LDS.128) ----
        const uint32_t jbase = static_cast<uint32_t>(p.t_index_base + tb + s * TS);
        sweep_dispatch<RMAX, RMAX, COL, TOP2, CS>(nr_w, reinterpret_cast<const uint4 *>(tile[b]), rows, jbase, q, b1,
                                                  b2, colbias, COL ? smem_u32(&colmin_s[b][warp][0]) : 0u, lane);
    }

    if (COL) {
        __syncthreads();
        // stages n_stage-2 and n_stage-1 still sit in shared memory
        for (int s = max(0, n_stage - 2); s < n_stage; ++s) flush_cols(s);
    }

#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        if (qrow[r] < nq) {
            const size_t row = static_cast<size_t>(p.row_out_off ? p.row_out_off[prob] : q_row0) + qrow[r];
            if (!TOP2 && p.compact) {
                uint32_t *c = reinterpret_cast<uint32_t *>(p.row_keys) + row;
                if (gridDim.y == 1)
                    *c = b1[r];
                else
                    atomicMin(c, b1[r]);
                continue;
            }
            uint2 *g = p.row_keys + row;
            if (gridDim.y == 1)
                *g = make_uint2(b1[r], b2[r]);
            else if (TOP2)
                merge_row_keys(g, b1[r], b2[r]);
            else
                atomicMin(&g->x, b1[r]);  // second key stays KEY_NONE (pre-set)
        }
    }
}

__global__ void unpack_keys_kernel(const uint32_t *__restrict__ keys, int64_t n, int32_t *__restrict__ idx,
                                   int32_t *__restrict__ dist)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    idx[i] = k == KEY_NONE ? -1 : static_cast<int32_t>(k & KEY_IDX_MASK);
    dist[i] = k == KEY_NONE ? -1 : static_cast<int32_t>(k >> KEY_IDX_BITS);
}

__global__ void merge_top2_kernel(const uint2 *__restrict__ shard, int n_shards, int nq, uint2 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
    for (int s = 0; s < n_shards; ++s) {
        const uint2 k = shard[static_cast<size_t>(s) * nq + i];
        const uint32_t hi = max(b1, k.x);
        b1 = min(b1, k.x);
        b2 = min(min(b2, hi), k.y);
    }
    out[i] = make_uint2(b1, b2);
}

__global__ void cross_check_kernel(const uint2 *__restrict__ row_keys, const uint32_t *__restrict__ col_keys, int nq,
                                   int nt, int32_t *__restrict__ match_t, int32_t *__restrict__ match_dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint32_t k = row_keys[i].x;
    int32_t mt = -1, md = -1;
    if (k != KEY_NONE) {
        const uint32_t j = k & KEY_IDX_MASK;
        if (j < static_cast<uint32_t>(nt)) {
            const uint32_t c = col_keys[j];
            if (c != KEY_NONE && (c & KEY_IDX_MASK) == static_cast<uint32_t>(i)) {
                mt = static_cast<int32_t>(j);
                md = static_cast<int32_t>(k >> KEY_IDX_BITS);
            }
        }
    }
    match_t[i] = mt;
    match_dist[i] = md;
}

__global__ void ratio_test_kernel(const uint2 *__restrict__ row_keys, int nq, int num, int den,
                                  uint8_t *__restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint2 k = row_keys[i];
    uint8_t m = 0;
    if (k.x != KEY_NONE && k.y != KEY_NONE) {
        const int d1 = static_cast<int>(k.x >> KEY_IDX_BITS), d2 = static_cast<int>(k.y >> KEY_IDX_BITS);
        m = (num * d1 < den * d2) ? 1 : 0;
    }
    mask[i] = m;
}

int gcd(int a, int b) { return b == 0 ? a : gcd(b, a % b); }

template <int THREADS, int RMAX, int CS>
int launch_hamming(const HammingParams &p, dim3 grid, bool top2, cudaStream_t stream)
{
    if (p.col_keys) {
        if (top2)
            hamming_top2_kernel<THREADS, RMAX, true, true, CS><<<grid, THREADS, 0, stream>>>(p);
        else
            hamming_top2_kernel<THREADS, RMAX, true, false, CS><<<grid, THREADS, 0, stream>>>(p);
    } else {
        if (top2)
            hamming_top2_kernel<THREADS, RMAX, false, true, CS><<<grid, THREADS, 0, stream>>>(p);
        else
            hamming_top2_kernel<THREADS, RMAX, false, false, CS><<<grid, THREADS, 0, stream>>>(p);
    }
    return launch_status();
}

// Shipping configuration (scripts/tune_matcher.py sweep on B200): 2 query rows per
// thread, 9 full adders per descriptor pair (the other variants of the sweep are in the git history).
constexpr int kRows = 2;
constexpr int kAdders = 9;

int big_rows() { return kRows; }
int launch_big(const HammingParams &p, dim3 grid, bool top2, cudaStream_t stream)
{
    return launch_hamming<256, kRows, kAdders>(p, grid, top2, stream);
}

// Kernel choice: SLAMFE_MATCH_MMA in `flags` selects the tcgen05 kernel (hamming_mma.cu) for this call;
// the environment variable SLAMFE_MATCH_MMA=1 / 0 (read once) forces it on / off for every call.
bool use_mma(int flags)
{
    static const int env = [] {
        const char *v = getenv("SLAMFE_MATCH_MMA");
        return v && *v ? (atoi(v) != 0 ? 1 : 0) : -1;
    }();
    if (env >= 0) return env == 1;
    return (flags & SLAMFE_MATCH_MMA) != 0;
}

// Pick the CTA shape and the train slicing so that the grid fills the SMs.
int run_hamming(HammingParams p, int n_problems, int max_nq, int max_nt, int64_t q_rows_total, int64_t t_rows_total,
                int flags, cudaStream_t stream)
{
    if (n_problems <= 0 || q_rows_total <= 0) return 0;
    const bool top2 = !(flags & SLAMFE_MATCH_BEST_ONLY);
    if ((flags & SLAMFE_MATCH_COMPACT_KEYS) && top2) return SLAMFE_EINVAL;
    p.compact = (flags & SLAMFE_MATCH_COMPACT_KEYS) ? 1 : 0;
    SLAMFE_CUDA_OK(cudaMemsetAsync(p.row_keys, 0xFF, (p.compact ? sizeof(uint32_t) : sizeof(uint2)) * q_rows_total,
                                   stream));
    if (p.col_keys && t_rows_total > 0)
        SLAMFE_CUDA_OK(cudaMemsetAsync(p.col_keys, 0xFF, sizeof(uint32_t) * t_rows_total, stream));
    if (max_nq <= 0 || max_nt <= 0) return 0;
    p.tma_quantum = 16 / gcd(p.t_stride, 16);
    if (use_mma(flags) && hamming_mma_supports(p.desc_bytes)) return run_hamming_mma(p, n_problems, max_nq, max_nt, top2, stream);

    const int sms = sm_count();
    const int stages_total = (max_nt + TS - 1) / TS;
    auto plan = [&](int tq, int min_stages_per_slice, int target, int &slices) {
        const long long ctas = static_cast<long long>((max_nq + tq - 1) / tq) * n_problems;
        slices = 1;
        if (ctas < target) {
            const int want = static_cast<int>((target + ctas - 1) / ctas);
            const int cap = max(1, stages_total / min_stages_per_slice);
            slices = min(want, cap);
        }
        return ctas * slices;
    };
    const int tq_big = 256 * big_rows();
    int slices_big = 1, slices_small = 1;
    const long long ctas_big = plan(tq_big, 4, 3 * sms, slices_big);
    const bool big = ctas_big >= 2LL * sms;
    if (!big) plan(128, 1, 4 * sms, slices_small);
    const int slices = big ? slices_big : slices_small;
    const int stages_per_slice = (stages_total + slices - 1) / slices;
    p.t_slice = stages_per_slice * TS;
    const int n_slices = (stages_total + stages_per_slice - 1) / stages_per_slice;
    const int tq = big ? tq_big : 128;
    const dim3 grid((max_nq + tq - 1) / tq, n_slices, n_problems);
    if (grid.y > 65535u || grid.z > 65535u) return SLAMFE_ERANGE;
    if (!big) return launch_hamming<128, 1, kAdders>(p, grid, top2, stream);
    return launch_big(p, grid, top2, stream);
}

int check_desc_args(const void *q, const void *t, int q_stride, int t_stride, int desc_bytes)
{
    if (!q || !t) return SLAMFE_EINVAL;
    if (desc_bytes <= 0 || desc_bytes > SLAMFE_MAX_DESC_BYTES) return SLAMFE_EINVAL;
    if (q_stride < desc_bytes || t_stride < desc_bytes || t_stride > SLAMFE_MAX_DESC_BYTES) return SLAMFE_EINVAL;
    return 0;
}

}  // namespace

}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_hamming_top2(const uint8_t *q, int nq, int q_stride, const uint8_t *t, int nt, int t_stride,
                                   int desc_bytes, int t_index_base, uint32_t *row_keys, uint32_t *col_keys,
                                   int flags, slamfe_stream_t stream)
{
    if (nq < 0 || nt < 0 || t_index_base < 0) return SLAMFE_EINVAL;
    if (nq == 0) return 0;
    if (!row_keys) return SLAMFE_EINVAL;
    if (nt > 0) {
        const int rc = check_desc_args(q, t, q_stride, t_stride, desc_bytes);
        if (rc) return rc;
    }
    if (static_cast<int64_t>(t_index_base) + nt > static_cast<int64_t>(KEY_IDX_MASK) ||
        nq > static_cast<int>(KEY_IDX_MASK))
        return SLAMFE_ERANGE;
    HammingParams p{};
    p.q = q; p.t = t; p.q_stride = q_stride; p.t_stride = t_stride; p.desc_bytes = desc_bytes;
    p.nq = nq; p.nt = nt; p.t_index_base = t_index_base;
    p.row_keys = reinterpret_cast<uint2 *>(row_keys);
    p.col_keys = col_keys;
    return run_hamming(p, 1, nq, nt, nq, nt, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int slamfe_hamming_top2_batched(const uint8_t *q, int q_stride, const int32_t *q_off, const int32_t *q_cnt,
                                           const uint8_t *t, int t_stride, const int32_t *t_off, const int32_t *t_cnt,
                                           int n_problems, int max_nq, int max_nt, int desc_bytes, int t_index_base,
                                           uint32_t *row_keys, int64_t q_rows_total, uint32_t *col_keys,
                                           int64_t t_rows_total, int flags, slamfe_stream_t stream)
{
    if (n_problems < 0 || max_nq < 0 || max_nt < 0 || q_rows_total < 0 || t_rows_total < 0 || t_index_base < 0)
        return SLAMFE_EINVAL;
    if (n_problems == 0 || q_rows_total == 0) return 0;
    if (!row_keys || !q_off || !t_off) return SLAMFE_EINVAL;
    const int rc = check_desc_args(q, t, q_stride, t_stride, desc_bytes);
    if (rc) return rc;
    if (max_nq > static_cast<int>(KEY_IDX_MASK) ||
        static_cast<int64_t>(t_index_base) + max_nt > static_cast<int64_t>(KEY_IDX_MASK))
        return SLAMFE_ERANGE;
    HammingParams p{};
    p.q = q; p.t = t; p.q_stride = q_stride; p.t_stride = t_stride; p.desc_bytes = desc_bytes;
    p.q_off = q_off; p.q_cnt = q_cnt; p.t_off = t_off; p.t_cnt = t_cnt; p.t_index_base = t_index_base;
    p.row_keys = reinterpret_cast<uint2 *>(row_keys);
    p.col_keys = col_keys;
    return run_hamming(p, n_problems, max_nq, max_nt, q_rows_total, t_rows_total, flags,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int slamfe_hamming_top2_pairs(const uint8_t *q, int q_stride, const int32_t *q_off, const int32_t *q_cnt,
                                         const uint8_t *t, int t_stride, const int32_t *t_off, const int32_t *t_cnt,
                                         const int32_t *out_off, int n_problems, int max_nq, int max_nt,
                                         int desc_bytes, uint32_t *row_keys, int64_t out_rows_total, int flags,
                                         slamfe_stream_t stream)
{
    if (n_problems < 0 || max_nq < 0 || max_nt < 0 || out_rows_total < 0) return SLAMFE_EINVAL;
    if (n_problems == 0 || out_rows_total == 0) return 0;
    if (!row_keys || !q_off || !t_off || !q_cnt || !t_cnt || !out_off) return SLAMFE_EINVAL;
    const int rc = check_desc_args(q, t, q_stride, t_stride, desc_bytes);
    if (rc) return rc;
    if (max_nq > static_cast<int>(KEY_IDX_MASK) || max_nt > static_cast<int>(KEY_IDX_MASK)) return SLAMFE_ERANGE;
    HammingParams p{};
    p.q = q; p.t = t; p.q_stride = q_stride; p.t_stride = t_stride; p.desc_bytes = desc_bytes;
    p.q_off = q_off; p.q_cnt = q_cnt; p.t_off = t_off; p.t_cnt = t_cnt; p.row_out_off = out_off;
    p.row_keys = reinterpret_cast<uint2 *>(row_keys);
    p.col_keys = nullptr;
    return run_hamming(p, n_problems, max_nq, max_nt, out_rows_total, 0, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int slamfe_unpack_keys(const uint32_t *keys, int64_t n, int32_t *idx, int32_t *dist, slamfe_stream_t stream)
{
    if (n < 0) return SLAMFE_EINVAL;
    if (n == 0) return 0;
    if (!keys || !idx || !dist) return SLAMFE_EINVAL;
    unpack_keys_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, n, idx,
                                                                                                            dist);
    return launch_status();
}

extern "C" int slamfe_merge_top2(const uint32_t *shard_keys, int n_shards, int nq, uint32_t *out, slamfe_stream_t stream)
{
    if (n_shards < 0 || nq < 0) return SLAMFE_EINVAL;
    if (nq == 0) return 0;
    if (!shard_keys || !out) return SLAMFE_EINVAL;
    merge_top2_kernel<<<(nq + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint2 *>(shard_keys), n_shards, nq, reinterpret_cast<uint2 *>(out));
    return launch_status();
}

extern "C" int slamfe_cross_check(const uint32_t *row_keys, const uint32_t *col_keys, int nq, int nt, int32_t *match_t,
                                  int32_t *match_dist, slamfe_stream_t stream)
{
    if (nq < 0 || nt < 0) return SLAMFE_EINVAL;
    if (nq == 0) return 0;
    if (!row_keys || !match_t || !match_dist || (nt > 0 && !col_keys)) return SLAMFE_EINVAL;
    cross_check_kernel<<<(nq + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint2 *>(row_keys), col_keys, nq, nt, match_t, match_dist);
    return launch_status();
}

extern "C" int slamfe_ratio_test(const uint32_t *row_keys, int nq, int num, int den, uint8_t *mask, slamfe_stream_t stream)
{
    if (nq < 0) return SLAMFE_EINVAL;
    if (nq == 0) return 0;
    if (!row_keys || !mask) return SLAMFE_EINVAL;
    ratio_test_kernel<<<(nq + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint2 *>(row_keys), nq, num, den, mask);
    return launch_status();
}
