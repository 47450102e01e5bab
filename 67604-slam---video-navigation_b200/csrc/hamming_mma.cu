// Brute-force Hamming matcher on the 5th-generation tensor cores (tcgen05, sm_100a) — same contract,
// same keys, same tie-breaks as hamming_top2_kernel in hamming.cu (reference call sites:
// final_project/backend/database/database.py:54-55, backend/loop/loop_closure.py:422,
// final_project/algorithms/matching.py:15,44, VAN_ex/code/ex1.py:189-190).
//
// Hamming distance as an exact integer contraction.  With query bits q_k and train bits t_k,
//     d(q, t) = sum_k q_k + sum_k t_k * (1 - 2 q_k) = popc(q) + A_q . B_t,
// A_q = (1 - 2 q_k) in {+1, -1} (int8), B_t = t_k in {0, 1} (uint8).  tcgen05.mma kind::i8 accumulates
// A . B^T exactly in int32, so distance, key = (d << 22) | index and therefore every tie-break are
// bit-identical to the XOR/POPC kernel.  The order of the 512 K positions is irrelevant as long as A
// and B agree, which makes the bit -> byte expansion cheap: output word s of input word w holds bits
// s, s+8, s+16, s+24 of w, i.e. (w >> s) & 0x01010101.
//
// One CTA = 128 query rows (UMMA M) against a train sweep in stages of 192 rows (UMMA N):
//   warps 0-3   epilogue: write the +-1 query tile into TMEM once (A operand, 128 columns), then per
//               stage tcgen05.ld the 128 x 192 int32 accumulators (lane = query row) and fold them into
//               the running row keys (1 IMAD + 1 add-min per pair); column minima (crossCheck /
//               backward match) by one redux.sync.min per column and warp, merged through shared
//               memory into one global atomicMin per train row and CTA;
//   warps 4-9   expanders: raw 61-byte train rows (1-D TMA bulk copy, double buffered) -> 0/1 bytes in
//               the K-major SWIZZLE_NONE core-matrix layout of the B operand (8 rows x 16 B contiguous,
//               K chunks LBO = 192*16 B apart), one row per lane, conflict-free STS.128;
//   warp 10     one thread issues the TMA copies and 16 tcgen05.mma (M128 N192 K32, A from TMEM) per
//               stage; tcgen05.commit hands the B stage back to the expanders and the accumulator
//               stage to the epilogue.  Two B stages (2 x 96 KB) and two accumulator stages
//               (2 x 192 TMEM columns) keep expansion, MMA and epilogue of consecutive stages overlapped.
// Facts pinned on a B200 by scripts/probe_tcgen05.cu (profiles/r02_probe_tcgen05.log): descriptor
// field meaning (LBO = K-chunk stride, SBO = 8-row stride), TMEM A layout (4 K-bytes per column),
// exactness, 135 cycles per M128 N256 K32 TS-mode MMA (94 % of the 128-cycle floor), TMEM reads
// >= 577 B/clk/SM, redux.sync ~1 per clk per SM.
#include "common.cuh"
#include "hamming_params.cuh"

namespace slamfe {

namespace {

constexpr int MQ = 128;                 // query rows per CTA = UMMA M
constexpr int NT = 192;                 // train rows per stage = UMMA N
constexpr int KCH = 32;                 // 16-byte K chunks per row (512 K positions)
constexpr int LBO = NT * 16;            // bytes between consecutive K chunks of the B tile
constexpr int B_STAGE = KCH * LBO;      // 98304
constexpr int RAW_STAGE = NT * SLAMFE_MAX_DESC_BYTES + 16;
constexpr int N_EPI_WARPS = 4, N_EXP_WARPS = 6;
constexpr int MMA_WARP = N_EPI_WARPS + N_EXP_WARPS;
constexpr int THREADS = (MMA_WARP + 1) * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_A = 2 * NT;     // columns 384..511 hold the query tile

struct __align__(16) Smem {
    uint8_t b[2][B_STAGE];
    uint8_t raw[2][RAW_STAGE];
    uint32_t colmin[2][N_EPI_WARPS][NT];
    uint64_t raw_full[2], b_full[2], b_empty[2], d_full[2], d_empty[2], a_ready;
    uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 B contiguous
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[tmem] . B[smem]^T, int8 x uint8 -> int32
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kIdesc = (2u << 4)                               // D format S32
                            | (1u << 7)                             // A signed 8-bit
                            | (0u << 10)                            // B unsigned 8-bit
                            | (static_cast<uint32_t>(NT >> 3) << 17)  // N
                            | (static_cast<uint32_t>(MQ >> 4) << 24); // M; both operands K-major

// bits s, s+8, s+16, s+24 of w as four 0/1 bytes
__device__ __forceinline__ uint32_t spread(uint32_t w, int s) { return (w >> s) & 0x01010101u; }

// Fold 32 accumulator columns [c0, c0+32) of one stage into the running keys.
template <bool COL, bool TOP2, bool MASKED>
__device__ __forceinline__ void fold_chunk(const uint32_t (&v)[32], int c0, int rows, uint32_t rowbase_j,
                                           uint32_t colbias, uint32_t &b1, uint32_t &b2, uint32_t colmin_addr, int lane)
{
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if (MASKED && c0 + c >= rows) break;
        // v = popc-free part of the distance: d = popc(q) + v; all arithmetic is exact mod 2^32
        const uint32_t key = v[c] * (1u << KEY_IDX_BITS) + rowbase_j + static_cast<uint32_t>(c);
        if (TOP2) {
            const uint32_t hi = max(b1, key);
            b2 = min(b2, hi);
        }
        b1 = min(b1, key);
        if (COL) {
            const uint32_t ck = v[c] * (1u << KEY_IDX_BITS) + colbias;
            const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, ck);
            if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(colmin_addr + 4u * (c0 + c)), "r"(m) : "memory");
        }
    }
}

// grid = (query tiles of 128 rows, train slices, problems)
template <bool COL, bool TOP2>
__global__ void __launch_bounds__(THREADS, 1) hamming_mma_kernel(const HammingParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

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
    const int qt0 = blockIdx.x * MQ;
    const int tb = blockIdx.y * p.t_slice;
    if (qt0 >= nq || tb >= nt) return;  // CTA-uniform; outputs were pre-set to KEY_NONE
    const int te = min(nt, tb + p.t_slice);
    const int n_stage = (te - tb + NT - 1) / NT;
    const int n_k = (p.desc_bytes + 3) >> 2;  // MMA K steps: 32 K positions = 4 descriptor bytes each

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.raw_full[i], 1);
            mbar_init(&sm.b_full[i], N_EXP_WARPS * 32);
            mbar_init(&sm.b_empty[i], 1);
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], N_EPI_WARPS * 32);
        }
        mbar_init(&sm.a_ready, N_EPI_WARPS * 32);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    const uint8_t *t_base = p.t + static_cast<size_t>(t_row0) * p.t_stride;
    auto stage_rows = [&](int s) { return min(NT, te - (tb + s * NT)); };
    auto stage_src = [&](int s) { return t_base + static_cast<size_t>(tb + s * NT) * p.t_stride; };
    auto stage_tma_rows = [&](int s) {
        if (reinterpret_cast<uintptr_t>(stage_src(s)) & 15) return 0;
        const int rows = stage_rows(s);
        return rows - rows % p.tma_quantum;
    };

    if (warp < N_EPI_WARPS) {
        // ============================== epilogue warps ==============================
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;  // this warp's TMEM lane quarter
        const int row = qt0 + tid;
        const int src_row = min(row, nq - 1);  // rows past the end mirror the last row (results not written)
        uint32_t pq = 0;
        {
            uint32_t w[W];
            load_desc_global(p.q + static_cast<size_t>(q_row0 + src_row) * p.q_stride, p.desc_bytes, w);
#pragma unroll
            for (int k = 0; k < W; ++k) {
                pq += __popc(w[k]);
                uint32_t a[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) a[s] = spread(w[k], s) * 0xFEu | 0x01010101u;  // bit 0 -> +1, bit 1 -> -1
                tmem_st8(tmem + lane_base + TMEM_A + 8 * k, a);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&sm.a_ready);
        }
        const uint32_t rowbase = pq << KEY_IDX_BITS;
        const uint32_t colbias = rowbase + static_cast<uint32_t>(src_row);
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        for (int s = 0; s < n_stage; ++s) {
            const int b = s & 1;
            const int rows = stage_rows(s);
            mbar_wait(&sm.d_full[b], (s >> 1) & 1);
            tc_fence_after();
            const uint32_t jstage = rowbase + static_cast<uint32_t>(p.t_index_base + tb + s * NT);
            const uint32_t colmin_addr = COL ? smem_u32(&sm.colmin[b][warp][0]) : 0u;
#pragma unroll 1
            for (int c0 = 0; c0 < rows; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_base + b * NT + c0, v);
                tmem_wait_ld();
                if (c0 + 32 <= rows)
                    fold_chunk<COL, TOP2, false>(v, c0, rows, jstage + c0, colbias, b1, b2, colmin_addr, lane);
                else
                    fold_chunk<COL, TOP2, true>(v, c0, rows, jstage + c0, colbias, b1, b2, colmin_addr, lane);
            }
            tc_fence_before();
            mbar_arrive(&sm.d_empty[b]);  // accumulator stage b may be overwritten
            if (COL) {
                // the 4 epilogue warps merge their column minima: one global atomicMin per train row
                asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");
                for (int c = tid; c < rows; c += N_EPI_WARPS * 32) {
                    const uint32_t m = min(min(sm.colmin[b][0][c], sm.colmin[b][1][c]),
                                           min(sm.colmin[b][2][c], sm.colmin[b][3][c]));
                    atomicMin(p.col_keys + t_row0 + tb + s * NT + c, m);
                }
                // colmin[b] is rewritten in stage s+2, after the bar.sync of stage s+1
            }
        }
        if (row < nq) {
            const size_t orow = static_cast<size_t>(p.row_out_off ? p.row_out_off[prob] : q_row0) + row;
            if (!TOP2 && p.compact) {
                uint32_t *c = reinterpret_cast<uint32_t *>(p.row_keys) + orow;
                if (gridDim.y == 1)
                    *c = b1;
                else
                    atomicMin(c, b1);
            } else {
                uint2 *g = p.row_keys + orow;
                if (gridDim.y == 1)
                    *g = make_uint2(b1, b2);
                else if (TOP2)
                    merge_row_keys(g, b1, b2);
                else
                    atomicMin(&g->x, b1);  // second key stays KEY_NONE (pre-set)
            }
        }
    } else if (warp < MMA_WARP) {
        // ============================== expander warps ==============================
        const int e = warp - N_EPI_WARPS;
        const int r = 32 * e + lane;  // this lane's train row within every stage
        const bool aligned = (p.t_stride & 3) == 0;
        for (int s = 0; s < n_stage; ++s) {
            const int b = s & 1;
            const int rows = stage_rows(s), trows = stage_tma_rows(s);
            if (trows > 0) mbar_wait(&sm.raw_full[b], (s >> 1) & 1);
            mbar_wait(&sm.b_empty[b], ((s >> 1) & 1) ^ 1);
            if (r < rows) {
                uint8_t *dst = sm.b[b] + r * 16;
                if (r < trows && aligned) {
                    // 4-byte aligned rows: lane l starts at word l (mod 16) so that both the LDS.32 of the 32
                    // rows (stride a multiple of 4 words) and the STS.128 spread over all banks
                    const uint32_t *row32 = reinterpret_cast<const uint32_t *>(sm.raw[b] + r * p.t_stride);
#pragma unroll 4
                    for (int i = 0; i < W; ++i) {
                        const int k = (i + lane) & (W - 1);
                        if (k < n_k) {
                            const uint32_t w = row32[k] & word_mask(p.desc_bytes, k);
                            uint8_t *d = dst + 2 * k * LBO;
                            *reinterpret_cast<uint4 *>(d) = make_uint4(spread(w, 0), spread(w, 1), spread(w, 2), spread(w, 3));
                            *reinterpret_cast<uint4 *>(d + LBO) =
                                make_uint4(spread(w, 4), spread(w, 5), spread(w, 6), spread(w, 7));
                        }
                    }
                } else {
                    uint32_t w[W];
                    if (r < trows) {  // unaligned rows (61-byte stride): 17 LDS.32 + funnel shifts
                        const int o = r * p.t_stride;
                        const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(sm.raw[b]) + (o >> 2);
                        const int sh = (o & 3) * 8;
                        uint32_t lo = raw32[0];
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            const uint32_t hi = raw32[k + 1];
                            w[k] = __funnelshift_r(lo, hi, sh) & word_mask(p.desc_bytes, k);
                            lo = hi;
                        }
                    } else {  // rows the bulk copy could not take (misaligned source / tail): read global memory
                        load_desc_global(stage_src(s) + static_cast<size_t>(r) * p.t_stride, p.desc_bytes, w);
                    }
#pragma unroll
                    for (int k = 0; k < W; ++k) {
                        if (k < n_k) {
                            uint8_t *d = dst + 2 * k * LBO;
                            *reinterpret_cast<uint4 *>(d) =
                                make_uint4(spread(w[k], 0), spread(w[k], 1), spread(w[k], 2), spread(w[k], 3));
                            *reinterpret_cast<uint4 *>(d + LBO) =
                                make_uint4(spread(w[k], 4), spread(w[k], 5), spread(w[k], 6), spread(w[k], 7));
                        }
                    }
                }
            }
            fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&sm.b_full[b]);
        }
    } else if (lane == 0) {
        // ============================== TMA + MMA issue (one thread) ==============================
        auto issue_tma = [&](int s) {
            const int trows = stage_tma_rows(s);
            if (trows > 0) {
                const uint32_t bytes = static_cast<uint32_t>(trows) * p.t_stride;
                mbar_arrive_expect_tx(&sm.raw_full[s & 1], bytes);
                tma_load_1d(sm.raw[s & 1], stage_src(s), bytes, &sm.raw_full[s & 1]);
            }
        };
        issue_tma(0);
        if (n_stage > 1) issue_tma(1);
        mbar_wait(&sm.a_ready, 0);
        tc_fence_after();
        for (int s = 0; s < n_stage; ++s) {
            const int b = s & 1;
            const uint32_t ph = (s >> 1) & 1;
            mbar_wait(&sm.b_full[b], ph);
            if (s + 2 < n_stage) issue_tma(s + 2);  // every expander has finished reading raw[b]
            mbar_wait(&sm.d_empty[b], ph ^ 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(sm.b[b]);
            for (int k = 0; k < n_k; ++k)
                umma_i8_ts(tmem + b * NT, tmem + TMEM_A + 8 * k, umma_desc(b_addr + 2 * k * LBO, LBO, 128), kIdesc,
                           k > 0);
            umma_commit(&sm.b_empty[b]);
            umma_commit(&sm.d_full[b]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <bool COL, bool TOP2>
int launch_mma(const HammingParams &p, dim3 grid, cudaStream_t stream)
{
    static bool configured = false;  // per instantiation; racing threads set the same value
    if (!configured) {
        SLAMFE_CUDA_OK(cudaFuncSetAttribute(hamming_mma_kernel<COL, TOP2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(sizeof(Smem))));
        configured = true;
    }
    hamming_mma_kernel<COL, TOP2><<<grid, THREADS, sizeof(Smem), stream>>>(p);
    return launch_status();
}

}  // namespace

int run_hamming_mma(HammingParams p, int n_problems, int max_nq, int max_nt, bool top2, cudaStream_t stream)
{
    // One CTA per SM (192 KB of shared memory, all 512 TMEM columns).  When the query tiles alone do not
    // fill the machine the train set is cut into slices (grid.y); their results merge exactly through
    // atomicMin / the CAS pair merge, as in the INT kernel.
    const int sms = sm_count();
    const int stages_total = (max_nt + NT - 1) / NT;
    const long long ctas = static_cast<long long>((max_nq + MQ - 1) / MQ) * n_problems;
    int slices = 1;
    if (ctas < 2LL * sms) {
        const int want = static_cast<int>((2LL * sms + ctas - 1) / ctas);
        slices = max(1, min(want, stages_total / 2));
    }
    const int stages_per_slice = (stages_total + slices - 1) / slices;
    p.t_slice = stages_per_slice * NT;
    const int n_slices = (stages_total + stages_per_slice - 1) / stages_per_slice;
    const dim3 grid((max_nq + MQ - 1) / MQ, n_slices, n_problems);
    if (grid.y > 65535u || grid.z > 65535u) return SLAMFE_ERANGE;
    if (p.col_keys) return top2 ? launch_mma<true, true>(p, grid, stream) : launch_mma<true, false>(p, grid, stream);
    return top2 ? launch_mma<false, true>(p, grid, stream) : launch_mma<false, false>(p, grid, stream);
}

}  // namespace slamfe
