// Development only (-DSLAMFE_MMA_DEV, SLAMFE_MMA_PAIR=1): measured and NOT shipped — identical keys on the whole shape
// matrix, but slower than the one-SM kernel wherever column minima are asked for (result at the end of this comment).
//
// Two-SM form of the tcgen05 matcher (hamming_mma.cu): a CLUSTER of two CTAs sweeps 256 query rows (128 per CTA)
// against the train rows with tcgen05.mma.cta_group::2 (M = 256, N = 128 train rows per stage, half of them expanded
// by each CTA).  Included by hamming_mma.cu inside namespace slamfe::{anonymous}; same contract, same keys, same
// tie-breaks as the one-SM kernel.
//
// Why.  With column minima (crossCheck / backward match: 90 % of a sequence step) the one-SM kernel is bound by shared
// memory bandwidth, not by the tensor pipe: per stage of 256 x 128 pairs (2048 cycles of MMA work) an SM moves
// 128 KB of B operand into the tensor core (the 64 KB stage is read once per query tile), 72 KB for the expansion
// (raw rows in, 0x80 bytes out) and 128 KB for the epilogue's transposes — 2.6 k cycles at 128 B/clk, 2.8 k measured
// (more ILP in the epilogue and fewer barrier arrivals changed nothing,
// profiles/r02_mma_scr2_ab.log, r02_persistent_ab.log).  In a CTA pair the hardware feeds both tensor cores from the
// two CTAs' half tiles: per 128 x 128 pairs (1024 cycles of MMA work) an SM reads 32 KB of B, spends 37 KB on the
// expansion of its 64 rows and 64 KB on the transposes — 1.07 k cycles, level with the tensor pipe.
// Facts pinned by scripts/probe_2cta.cu on a B200 (profiles/r02_probe_2cta.log): A in TMEM works with
// cta_group::2 (each CTA holds its own 128 rows at the same TMEM address), CTA r's shared-memory tile supplies
// accumulator columns [r N/2, (r + 1) N/2), one thread of the leader issues, tcgen05.commit ... multicast::cluster
// arrives on the barrier at the same offset in both CTAs, a remote mbarrier.arrive.release.cluster from the peer is seen
// by the leader's try_wait.acquire.cluster, and M256 N128 K32 takes 64 cycles: 16 descriptor pairs/clk/SM, the same floor.
//
// A first version kept the one-SM kernel's shape (two query tiles and two accumulators per CTA): identical keys, 0.7 x
// the speed (profiles/r02_pair_v1_ab.log).  Draining an accumulator takes as long as the MMA that fills it (64 KB at
// the 64 B/clk of the TMEM read path = 1024 cycles), so with two accumulators the loop has no slack, and the hand-offs
// of a pair (multicast commit -> epilogue, cluster-scope arrival -> issuing lane) went straight into it.  Hence ONE
// query tile per CTA and THREE accumulators: the +-1 tile takes 128 TMEM columns, the accumulators 3 x 128.
//
// Roles per CTA (14 warps):
//   warps 0-7   epilogue: warp w owns TMEM lane quarter w & 3 (32 query rows); warps 0-3 fold the even stages, warps
//               4-7 the odd ones (both 64-column chunks of the stage's accumulator: tcgen05.ld, row minima, column
//               minima through the swizzled transpose, exactly as a query tile's warps in the one-SM kernel — a split
//               by column chunk with every warp in every stage was 2 x slower, the per-stage chain of waits and
//               barriers did not fit 1024 cycles); the two warps of a quarter merge their row keys through the global
//               atomicMin / CAS merge the sliced launches use; the four warps of a group merge column minima in
//               shared memory (named barrier 1 + group), one global atomicMin per train row and stage;
//   warps 8-11  expanders: warps 8-9 expand this CTA's 64 rows of the even stages, warps 10-11 those of the odd
//               stages (a thread needs ~1300 cycles per row, a stage lasts 1024);
//   warp 12     MMA issue, one lane, LEADER CTA ONLY: waits on its own barriers, which count one elected arrival per
//               warp of BOTH CTAs, issues 16 x M256 N128 K32 per stage, multicasts its commits to both CTAs;
//   warp 13     TMA lane: this CTA's 64 raw rows per stage, four staging buffers.
//
// Result on a B200 (profiles/r02_pair_ab.log; one-SM kernel in brackets): dense 20 k x 20 k rows only 0.163 ms
// (0.153), with column minima 0.252 ms (0.172); ragged stereo launch of 256 frames with column minima 1.58 ms (1.01).
// Without column minima the pair reaches the one-SM kernel's tensor-pipe utilisation (ncu: 80 %); with them nothing is
// saturated (ncu: L1/TEX 51 %, TMEM 54 %, issue slots 38 %, 61 % of cycles without an eligible warp) and timing-only
// variants put the loss on the shared-memory work of the column minima (no transposes: 0.166 ms; no per-stage flush:
// 0.210 ms): the operand sharing of a CTA pair did not relieve this SM's shared memory the way the B-traffic
// arithmetic above assumes, and a pair runs in lock-step (every stage waits for the slower CTA's expanders and
// epilogue).  The one-SM kernel stays the product.
#pragma once

struct GeoPair {
    static constexpr int NT = 128, NH = 64;             // train rows per stage: of the pair / expanded by this CTA
    static constexpr int CQ = MQ;                       // query rows per CTA (the pair sweeps 2 CQ)
    static constexpr int LBO = NH * 16;                 // bytes between consecutive K chunks of this CTA's half tile
    static constexpr int B_STAGE = KCH * LBO;
    static constexpr int NB = 4;                        // B (and raw) stages
    static constexpr int NACC = 3;                      // accumulators of NT columns
    static constexpr int RAW_STAGE = NH * SLAMFE_MAX_DESC_BYTES + 16;
    static constexpr int N_EPI_WARPS = 8, N_EXP_WARPS = 4;
    static constexpr int MMA_WARP = N_EPI_WARPS + N_EXP_WARPS;
    static constexpr int TMA_WARP = MMA_WARP + 1;
    static constexpr int THREADS = (TMA_WARP + 1) * 32;
    static constexpr uint32_t TMEM_A = NACC * NT;       // first column of the +-1 tile
    static constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (0u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                      (static_cast<uint32_t>((2 * MQ) >> 4) << 24);   // S32 = s8 x u8, N = 128, M = 256
    static_assert(NACC * NT + 128 <= 512, "TMEM: the accumulators + the query tile");
    struct __align__(16) Smem {
        uint8_t b[NB][B_STAGE];
        uint8_t raw[NB][RAW_STAGE];
        uint32_t scratch[N_EPI_WARPS][32 * 32];   // per-warp 32 x 64 packed distances, swizzled
        uint32_t colmin[2][2][NT];                // per epilogue group, double buffered over its stages
        uint64_t raw_full[NB], raw_empty[NB], b_full[NB], b_empty[NB], d_full[NACC], d_empty[NACC], a_ready;
        uint32_t tmem_base;
    };
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER's copy of a barrier (CTA 0 of the pair), from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)
{
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(raddr) : "r"(smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait on a local barrier whose arrivals come from the peer CTA or from a multicast tcgen05.commit
__device__ __forceinline__ bool mbar_try_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_cluster(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_wait_relaxed_cluster(uint64_t *bar, uint32_t parity, unsigned ns)
{
    while (!mbar_try_cluster(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// D[tmem, both CTAs] (+)= A[tmem, both CTAs] . B[smem halves of both CTAs]^T, int8 x uint8 -> int32
__device__ __forceinline__ void umma_i8_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// grid = (2 x query blocks of 256 rows, train slices, problems), clusters of two CTAs along x
template <bool COL, bool TOP2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GeoPair::THREADS, 1) hamming_mma_pair_kernel(const HammingParams p)
{
    using G = GeoPair;
    constexpr int NT = G::NT, NH = G::NH, CQ = G::CQ, LBO = G::LBO, NB = G::NB, NACC = G::NACC;
    constexpr int N_EPI_WARPS = G::N_EPI_WARPS, MMA_WARP = G::MMA_WARP, THREADS = G::THREADS;
    using Smem = typename G::Smem;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int prob = blockIdx.z;
    const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs for the pair)

    int q_row0 = 0, nq = p.nq, t_row0 = 0, nt = p.nt;
    if (p.q_off) {
        q_row0 = p.q_off[prob];
        nq = p.q_cnt ? p.q_cnt[prob] : p.q_off[prob + 1] - q_row0;
    }
    if (p.t_off) {
        t_row0 = p.t_off[prob];
        nt = p.t_cnt ? p.t_cnt[prob] : p.t_off[prob + 1] - t_row0;
    }
    const int qp0 = static_cast<int>(blockIdx.x >> 1) * (2 * CQ);   // first query row of the pair
    const int tb = blockIdx.y * p.t_slice;
    if (qp0 >= nq || tb >= nt) return;  // cluster-uniform; outputs were pre-set to KEY_NONE
    // This CTA's 128 query rows.  The M = 256 MMA always computes both CTAs' halves, so a CTA whose rows lie past the
    // end of the problem still runs the whole protocol (its rows mirror row nq - 1; nothing of it is written out, and
    // its column minima are skipped).
    const int qt0 = qp0 + static_cast<int>(rank) * CQ;
    const bool has_rows = qt0 < nq;
    const int te = min(nt, tb + p.t_slice);
    const int n_stage = (te - tb + NT - 1) / NT;
    const int ws = p.desc_bytes >> 2;   // input word whose byte 3 holds the 8 spare K positions
    const int n_k = ws + 1;             // MMA K steps: 32 K positions = 4 descriptor bytes each

    // epilogue threads: start fetching their query row now, its HBM latency overlaps the TMEM allocation and
    // the barrier set-up below (the two warps of a lane quarter fetch the same rows: each builds half of the K range)
    uint32_t wq[W];
    if (warp < N_EPI_WARPS) {
        const int row_q = qt0 + (warp & 3) * 32 + lane;
        load_desc_global(p.q + static_cast<size_t>(q_row0 + min(row_q, nq - 1)) * p.q_stride, p.desc_bytes, wq);
    }
    if (tid == 0) {
        // b_full, d_empty and a_ready are used in the LEADER only and count one elected arrival per warp of BOTH
        // CTAs; raw_full / raw_empty are local; b_empty and d_full receive the multicast commit in each CTA
        for (int i = 0; i < NB; ++i) {
            mbar_init(&sm.raw_full[i], 1);
            mbar_init(&sm.raw_empty[i], 2);
            mbar_init(&sm.b_full[i], 2 * 2);
            mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], 2 * 4);
        }
        mbar_init(&sm.a_ready, 2 * N_EPI_WARPS);
        fence_mbar_init();
    }
    if (COL)
        for (int i = tid; i < 2 * 2 * NT; i += THREADS) (&sm.colmin[0][0][0])[i] = KEY_NONE;
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    const uint8_t *t_base = p.t + static_cast<size_t>(t_row0) * p.t_stride;
    auto stage_rows = [&](int s) { return min(NT, te - (tb + s * NT)); };   // valid rows of the whole stage
    auto stage_src = [&](int s) {    // this CTA's half: stage rows [rank * NH, rank * NH + NH)
        return t_base + static_cast<size_t>(tb + s * NT + static_cast<int>(rank) * NH) * p.t_stride;
    };
    auto half_rows = [&](int s) { return max(0, min(NH, stage_rows(s) - static_cast<int>(rank) * NH)); };
    auto stage_tma_rows = [&](int s) {   // rows of this CTA's half the bulk copy takes
        if (reinterpret_cast<uintptr_t>(stage_src(s)) & 15) return 0;
        return half_rows(s) & ~(p.tma_quantum - 1);   // the quantum is a power of two (16 / gcd(stride, 16))
    };

    if (warp < N_EPI_WARPS) {
        // ============================== epilogue warps ==============================
        const int quarter = warp & 3;   // TMEM lane quarter = 32 query rows
        const int g = warp >> 2;        // warps 0-3 fold the even stages, warps 4-7 the odd ones; K half of the +-1 tile
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;  // this warp's TMEM lanes
        const int row_in_tile = quarter * 32 + lane;
        const int row = qt0 + row_in_tile;
        // rows past the end mirror the last row (min(row, nq - 1) at kernel entry): their results are not
        // written, and for column minima they lose every tie to the real row
        const uint32_t one = static_cast<uint32_t>(p.desc_bytes > 0);  // opaque 1: keeps key adds on the FMA pipe
        {
            // the +-1 tile (see hamming_mma_kernel); warp g of a quarter writes K steps [8 g, 8 g + 8)
            uint32_t pq = 0;
#pragma unroll
            for (int k = 0; k < W; ++k) pq += __popc(wq[k]);
            uint32_t spare[8];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const uint32_t c = min(pq, 127u);
                spare[s] = c << 24;
                pq -= c;
            }
#pragma unroll
            for (int s = 4; s < 8; ++s) spare[s] = 127u << 24;
            uint32_t mul[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) mul[s] = (1u << (7 - s)) * one;
#pragma unroll
            for (int k = 0; k < W; ++k) {
                if (k < n_k && (k >> 3) == g) {
                    uint32_t a[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        uint32_t t, m;
                        asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(t) : "r"(wq[k]), "r"(mul[s]));
                        asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(m) : "r"(t));   // sign-replicate every byte
                        a[s] = m | 0x01010101u;                                    // bit 0 -> +1, bit 1 -> -1
                    }
                    if (k == ws) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) a[s] = (a[s] & 0x00FFFFFFu) | spare[s];
                    }
                    tmem_st8(tmem + lane_base + G::TMEM_A + 8 * k, a);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&sm.a_ready);
        }
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        const uint32_t scr_addr = smem_u32(sm.scratch[warp]);
        const uint32_t rbase = static_cast<uint32_t>(qt0 + quarter * 32);
        // group g folds the stages s = g, g + 2, ...: both 64-column chunks of the stage's accumulator, 2048 cycles of
        // MMA work apart — the same rhythm per warp as in the one-SM kernel
        int acc = g;          // s % NACC
        uint32_t use = 0;     // s / NACC
        for (int s = g; s < n_stage; s += 2) {
            const int rows = stage_rows(s);
            const int cm = (s >> 1) & 1;   // column-minima buffer of this group's stage
            mbar_wait_relaxed_cluster(&sm.d_full[acc], use & 1, 20);
            tc_fence_after();
            const uint32_t jstage = static_cast<uint32_t>(p.t_index_base + tb + s * NT);
            // read every 64-column chunk that holds a valid train row, then hand the accumulator back to
            // the tensor core BEFORE folding
            uint32_t vv[2][32];
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (h == 0 || 64 * h < rows) tmem_ld64_packed(tmem + lane_base + acc * NT + 64 * h, vv[h]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&sm.d_empty[acc]);
            acc += 2;
            if (acc >= NACC) {
                acc -= NACC;
                ++use;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h > 0 && 64 * h >= rows) break;  // warp-uniform
                uint32_t (&v)[32] = vv[h];
                if (COL && has_rows) {
                    // scratch[row = lane][32 words], 16-byte chunk i stored at chunk i ^ (lane & 7):
                    // conflict-free both for these row-wise STS.128 and for the column-wise LDS.32 below
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t a = scr_addr + lane * 128 + ((i ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[4 * i]), "r"(v[4 * i + 1]),
                                     "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                                     : "memory");
                    }
                }
                // ---- row minima: key16 = acc + column = (d << 7) | column-in-chunk ----
                const uint32_t jh = jstage + 64u * h;
                if (!TOP2) {
                    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint32_t k0 = add_imad(v[j], one, ((2u * j + 1u) << 16) | (2u * j));
                        const uint32_t k1 = add_imad(v[j + 1], one, ((2u * j + 3u) << 16) | (2u * j + 2u));
                        m = __vimin3_u16x2(m, k0, k1);
                    }
                    const uint32_t k16 = min(m & 0xFFFFu, m >> 16);
                    b1 = min(b1, ((k16 >> 7) << KEY_IDX_BITS) + jh + (k16 & 127u));
                } else {
                    uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;  // per 16-bit half: best and second best
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t k = add_imad(v[j], one, ((2u * j + 1u) << 16) | (2u * j));
                        m2 = __vminu2(m2, __vmaxu2(m1, k));
                        m1 = __vminu2(m1, k);
                    }
                    const uint32_t c[4] = {m1 & 0xFFFFu, m1 >> 16, m2 & 0xFFFFu, m2 >> 16};
#pragma unroll
                    for (int i = 0; i < 4; ++i)  // empty slots and invalid train rows (d = 508) are no candidates
                        top2_insert(c[i] >= (505u << 7) ? KEY_NONE : ((c[i] >> 7) << KEY_IDX_BITS) + jh + (c[i] & 127u), b1, b2);
                }
                if (COL && has_rows) {
                    // ---- column minima over this warp's 32 query rows; lane l ends up with the packed
                    //      minima of train columns 2l, 2l+1 of the chunk: key16 = (d << 5) | row-in-warp ----
                    uint32_t m = 0xFFFFFFFFu;
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 32; r += 2) {
                        uint32_t x0, x1;
                        const uint32_t a0 = scr_addr + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2));
                        const uint32_t a1 = scr_addr + (r + 1) * 128 + ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2));
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(a0) : "memory");
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x1) : "r"(a1) : "memory");
                        const uint32_t k0 = add_imad(x0 >> 2, one, (static_cast<uint32_t>(r) << 16) | r);
                        const uint32_t k1 = add_imad(x1 >> 2, one, (static_cast<uint32_t>(r + 1) << 16) | (r + 1));
                        m = __vimin3_u16x2(m, k0, k1);
                    }
                    __syncwarp();  // the scratch is rewritten by the next chunk / stage
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const uint32_t k16 = hh ? (m >> 16) : (m & 0xFFFFu);
                        const int c = 64 * h + 2 * lane + hh;
                        if (c < rows) atomicMin(&sm.colmin[g][cm][c], ((k16 >> 5) << KEY_IDX_BITS) + rbase + (k16 & 31u));
                    }
                }
            }
            if (COL && has_rows) {
                // the 4 warps of this group merge: one global atomicMin per train row and stage
                if (g == 0)
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                else
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                for (int c = row_in_tile; c < rows; c += 128) {
                    atomicMin(p.col_keys + t_row0 + tb + s * NT + c, sm.colmin[g][cm][c]);
                    sm.colmin[g][cm][c] = KEY_NONE;  // reused two of this group's stages on, after its next barrier
                }
            }
        }
        if (row < nq) {
            // two warps (even / odd stages) hold results for this row: merge through global memory (pre-set to KEY_NONE)
            const size_t orow = static_cast<size_t>(p.row_out_off ? p.row_out_off[prob] : q_row0) + row;
            if (!TOP2 && p.compact) {
                atomicMin(reinterpret_cast<uint32_t *>(p.row_keys) + orow, b1);
            } else {
                uint2 *g = p.row_keys + orow;
                if (TOP2)
                    merge_row_keys(g, b1, b2);
                else
                    atomicMin(&g->x, b1);  // second key stays KEY_NONE (pre-set)
            }
        }
    } else if (warp < MMA_WARP) {
        // ============================== expander warps ==============================
        const int ew = warp - N_EPI_WARPS;
        const int eg = ew >> 1;                 // warps 8-9: even stages, warps 10-11: odd stages
        const int r = (ew & 1) * 32 + lane;     // this thread's row within this CTA's half of its stages
        uint32_t mul[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) mul[s] = (1u << (7 - s)) * static_cast<uint32_t>(p.desc_bytes > 0);
        const uint32_t last_mask = word_mask(p.desc_bytes, ws);   // the word that holds the descriptor's tail + the spare byte
        // Two copies of the stage loop, chosen once per kernel (see hamming_mma_kernel)
        auto sweep = [&](auto rows64_c) {
            constexpr bool ROWS64 = decltype(rows64_c)::value;
            for (int s = eg; s < n_stage; s += 2) {
                const int b = s % NB;
                const uint32_t ph = (s / NB) & 1;
                // valid rows of this CTA's half, and those of them that came through the bulk copy
                const int rows = half_rows(s), trows = stage_tma_rows(s);
                if (trows > 0) mbar_wait(&sm.raw_full[b], ph);
                mbar_wait_cluster(&sm.b_empty[b], ph ^ 1);   // the pair's multicast commit
                const bool fast64 = ROWS64 && r < trows && r < rows;
                if (fast64) {
                    // 64-byte rows: conflict-free 16-byte chunk reads, words straight to their K positions
                    uint8_t *dst64 = sm.b[b] + r * 16;
                    const uint4 *raw128 = reinterpret_cast<const uint4 *>(sm.raw[b] + r * 64);
                    const uint32_t slo = 0x80000000u;   // a valid row: 0x80 on the four popc positions of the tail word
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cc = (c + (r >> 1)) & 3;
                        const uint4 v = raw128[cc];
                        const bool tail = cc == 3;
                        const uint32_t ww[4] = {v.x, v.y, v.z, tail ? (v.w & last_mask) : v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t sl = (tail && j == 3) ? slo : 0u;
                            uint8_t *d = dst64 + 2 * (4 * cc + j) * LBO;
                            *reinterpret_cast<uint4 *>(d) = make_uint4(spread80(ww[j], mul[0]) | sl, spread80(ww[j], mul[1]) | sl,
                                                                       spread80(ww[j], mul[2]) | sl, spread80(ww[j], mul[3]) | sl);
                            *reinterpret_cast<uint4 *>(d + LBO) = make_uint4(spread80(ww[j], mul[4]), spread80(ww[j], mul[5]),
                                                                             spread80(ww[j], mul[6]), spread80(ww[j], mul[7]));
                        }
                    }
                } else {
                    uint32_t w[W];
                    if (r >= rows) {
#pragma unroll
                        for (int k = 0; k < W; ++k) w[k] = 0;
                    } else if (r < trows) {
                        const int o = r * p.t_stride;  // any alignment: LDS.32 + funnel shift
                        const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(sm.raw[b]) + (o >> 2);
                        const int sh = (o & 3) * 8;
                        uint32_t lo = raw32[0];
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            // words past the descriptor hold stale bytes of the staging buffer (never past its end:
                            // RAW_STAGE has 16 bytes of slack); they are masked / skipped below
                            const uint32_t hi = raw32[k + 1];
                            w[k] = __funnelshift_r(lo, hi, sh);
                            lo = hi;
                        }
                    } else {  // rows the bulk copy could not take (misaligned source / tail): read global memory
                        load_desc_global(stage_src(s) + static_cast<size_t>(r) * p.t_stride, p.desc_bytes, w);
                    }
                    uint8_t *dst = sm.b[b] + r * 16;
                    // valid rows carry 0x80 on the four popc positions, invalid rows on the four marker positions
                    const uint32_t spare_lo = (r < rows) ? 0x80000000u : 0u, spare_hi = (r < rows) ? 0u : 0x80000000u;
                    auto put = [&](int kk, uint32_t wk, uint32_t slo, uint32_t shi) {
                        uint8_t *d = dst + 2 * kk * LBO;
                        *reinterpret_cast<uint4 *>(d) = make_uint4(spread80(wk, mul[0]) | slo, spread80(wk, mul[1]) | slo,
                                                                   spread80(wk, mul[2]) | slo, spread80(wk, mul[3]) | slo);
                        *reinterpret_cast<uint4 *>(d + LBO) = make_uint4(spread80(wk, mul[4]) | shi, spread80(wk, mul[5]) | shi,
                                                                         spread80(wk, mul[6]) | shi, spread80(wk, mul[7]) | shi);
                    };
                    if (n_k == W) {   // descriptors of 60..63 bytes (AKAZE: 61): word 15 is the tail word, no per-word tests
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            if (k == W - 1)
                                put(k, w[k] & last_mask, spare_lo, spare_hi);
                            else
                                put(k, w[k], 0u, 0u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            if (k < ws)
                                put(k, w[k], 0u, 0u);
                            else if (k == ws)
                                put(k, w[k] & last_mask, spare_lo, spare_hi);
                        }
                    }
                }
                fence_async_smem();  // generic-proxy stores -> visible to the tensor cores of the pair (async proxy)
                __syncwarp();        // ... of every lane, before the warp's one arrival
                if (lane == 0) {
                    mbar_arrive(&sm.raw_empty[b]);          // local: the staging buffer may take its next bulk copy
                    mbar_arrive_leader(&sm.b_full[b]);      // release.cluster: this warp's rows of the stage are in place
                }
            }
        };
        if (n_k == W && p.t_stride == 64)
            sweep(std::true_type{});
        else
            sweep(std::false_type{});
    } else if (warp == MMA_WARP) {
        // ============================== MMA issue (one elected lane of the leader) ===========
        if (rank == 0 && elect_one()) {
            mbar_wait_cluster(&sm.a_ready, 0);
            tc_fence_after();
            // K-step k of a B stage: descriptor start address advances by 2 chunks = 2 * LBO bytes
            const uint64_t desc0 = umma_desc(smem_u32(sm.b[0]), LBO, 128);
            // the barriers of stage s + 1 are waited for in the MIDDLE of issuing stage s (see hamming_mma_kernel)
            int acc_n = 0;        // (s + 1) % NACC
            uint32_t use_n = 0;   // (s + 1) / NACC
            auto wait_stage = [&](int s) {   // its B stage (both halves) and its accumulator (drained by both CTAs)
                mbar_wait_cluster(&sm.b_full[s % NB], (s / NB) & 1);
                mbar_wait_cluster(&sm.d_empty[acc_n], (use_n & 1) ^ 1);
                tc_fence_after();
            };
            wait_stage(0);
            for (int s = 0; s < n_stage; ++s) {
                const int b = s % NB, acc = acc_n;
                if (++acc_n == NACC) {
                    acc_n = 0;
                    ++use_n;
                }
                const uint64_t desc_b = desc0 + static_cast<uint64_t>(b * (G::B_STAGE >> 4));
                const uint32_t d_addr = tmem + acc * NT, a_addr = tmem + G::TMEM_A;
                if (n_k == 16) {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        umma_i8_ts_pair(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    if (s + 1 < n_stage) wait_stage(s + 1);
#pragma unroll
                    for (int k = 12; k < 16; ++k)
                        umma_i8_ts_pair(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, 1);
                } else {
                    for (int k = 0; k < n_k; ++k)
                        umma_i8_ts_pair(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    if (s + 1 < n_stage) wait_stage(s + 1);
                }
                umma_commit_pair(&sm.b_empty[b]);
                umma_commit_pair(&sm.d_full[acc]);
            }
        }
    } else {
        // ============================== TMA producer (one elected lane) ===========
        // raw[b] is free again once both expander warps of its stages have arrived on raw_empty[b]
        if (elect_one()) {
            for (int s = 0; s < n_stage; ++s) {
                if (s >= NB) mbar_wait_relaxed(&sm.raw_empty[s % NB], ((s - NB) / NB) & 1, 256);
                const int trows = stage_tma_rows(s);
                if (trows > 0) {
                    const uint32_t bytes = static_cast<uint32_t>(trows) * p.t_stride;
                    mbar_arrive_expect_tx(&sm.raw_full[s % NB], bytes);
                    tma_load_1d(sm.raw[s % NB], stage_src(s), bytes, &sm.raw_full[s % NB]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody arrives on a barrier of a CTA that has left, and both halves of TMEM are idle
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <bool COL, bool TOP2>
int launch_mma_pair(const HammingParams &p, dim3 grid, cudaStream_t stream)
{
    static bool configured = false;  // per instantiation; racing threads set the same value
    constexpr int smem = static_cast<int>(sizeof(GeoPair::Smem));
    static_assert(smem <= 227 * 1024, "shared memory per CTA");
    if (!configured) {
        SLAMFE_CUDA_OK(cudaFuncSetAttribute(hamming_mma_pair_kernel<COL, TOP2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    hamming_mma_pair_kernel<COL, TOP2><<<grid, GeoPair::THREADS, smem, stream>>>(p);
    return launch_status();
}

// grid.x = 2 CTAs per block of 256 query rows; the train set is cut into slices when the pairs alone do not fill the SMs
int run_pair(HammingParams p, int n_problems, int max_nq, int max_nt, bool top2, cudaStream_t stream)
{
    using G = GeoPair;
    const int sms = sm_count();
    const int stages_total = (max_nt + G::NT - 1) / G::NT;
    const int blocks = (max_nq + 2 * G::CQ - 1) / (2 * G::CQ);
    const long long ctas = 2LL * blocks * n_problems;
    int slices = 1;
    if (ctas < 2LL * sms) {
        const int want = static_cast<int>((2LL * sms + ctas - 1) / ctas);
        slices = max(1, min(want, stages_total / 8));
    }
    const int stages_per_slice = (stages_total + slices - 1) / slices;
    p.t_slice = stages_per_slice * G::NT;
    const int n_slices = (stages_total + stages_per_slice - 1) / stages_per_slice;
    const dim3 grid(2 * blocks, n_slices, n_problems);
    if (grid.y > 65535u || grid.z > 65535u) return SLAMFE_ERANGE;
    if (p.col_keys) return top2 ? launch_mma_pair<true, true>(p, grid, stream) : launch_mma_pair<true, false>(p, grid, stream);
    return top2 ? launch_mma_pair<false, true>(p, grid, stream) : launch_mma_pair<false, false>(p, grid, stream);
}
