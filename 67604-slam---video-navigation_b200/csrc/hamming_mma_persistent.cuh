// Development only (-DSLAMFE_MMA_DEV): the PERSISTENT form of hamming_mma_kernel, measured and not shipped.
// Included by hamming_mma.cu inside namespace slamfe::{anonymous}, after the one-job-per-CTA kernel whose helpers,
// geometry and shared-memory layout it reuses.  SLAMFE_MMA_PERSISTENT=1 selects it in a development build
// (scripts/gpu_persist_ab.sh); tests/test_mma_persistent_protocol.py models its barrier protocol on the host.
// Result on a B200 (profiles/r02_persistent_ab.log): identical keys on the whole shape matrix and the GPU suite, and
// the same speed as the shipped kernel to within 1 % on the ragged launches of the sequence workload (-5 % on the
// single-problem dense sweep with column minima) — the per-CTA prologue it removes was not what bounds the launch
// ("What bounds it").
#pragma once

// ---------------------------------------------------------------------------------------------------
// Persistent form of the kernel above: one CTA per SM walks a dynamic list of (query tile, train slice,
// problem) jobs, so everything a one-job CTA pays before its first MMA (parameter loads, barrier set-up,
// TMEM allocation, the HBM latency of the query rows: ~7.6 k cycles, 11 % of a 28-stage job of the ragged
// stereo launch and 21 % of a 14-stage job of the pair stage) is paid once per SM, and the only per-job
// cost left on the tensor pipe is rebuilding the +-1 query tile in TMEM — which tile 0's warps do while
// the tensor core finishes tile 1 of the old job, and tile 1's warps while it starts tile 0 of the new one.
//   * jobs come from an atomic counter (p.job_counter, zero between launches: the last fetch of a launch
//     resets it), x fastest so that the CTAs running at the same time share a problem's train rows in L2;
//     the TMA lane fetches job n + 1 while job n streams in, decodes it, skips empty tiles of the ragged
//     job space and publishes the descriptor through a 4-deep shared-memory ring (sched_full / sched_empty);
//   * every barrier keeps running across jobs: each role counts B stages, accumulator uses, armed raw
//     buffers and A-tile builds on its own, identically;
//   * the epilogue warps fetch the NEXT job's query row at the start of a job (its latency hides behind the
//     sweep) and write the next +-1 tile right after the last accumulator of the current job has been read.
#ifdef SLAMFE_MMA_PROF
__device__ unsigned long long g_mma_prof[256][16];   // development only: clock64 sums per CTA (scripts/prof_persist.py)
#define PROF_NOW() clock64()
#define PROF_ADD(var, t0) (var) += clock64() - (t0)
#else
#define PROF_NOW() 0ll
#define PROF_ADD(var, t0) ((void)(t0))
#endif
constexpr int JR = 4;   // job ring depth
struct __align__(16) JobInfo {
    int valid, prob, q_row0, nq;   // valid = 0: no more jobs
    int qt0, t_row0, tb, te;
    int n_stage, n_tiles, out_row0, pad;
};
template <class G>
struct __align__(16) PSmem {
    typename G::Smem s;
    JobInfo jobs[JR];
    uint32_t qstage[G::N_EPI_WARPS * 32][17];   // next job's query rows, one 17-word slot per epilogue thread (cp.async)
    uint64_t sched_full[JR], sched_empty[JR], a_ready[2], raw_empty[2];
};

template <class G, bool COL, bool TOP2>
__global__ void __launch_bounds__(G::THREADS, 1) hamming_mma_persistent_kernel(const HammingParams p)
{
    constexpr int QT = G::QT, NT = G::NT, CQ = G::CQ, LBO = G::LBO, NB = G::NB;
    constexpr int N_EPI_WARPS = G::N_EPI_WARPS, MMA_WARP = G::MMA_WARP, THREADS = G::THREADS;
    // Every multi-thread barrier counts WARPS: the lanes of a warp order themselves with __syncwarp and one lane arrives
    // (128 single-lane arrivals on one mbarrier word are 128 serialised shared-memory atomics per stage and barrier).
    constexpr uint32_t N_READERS = N_EPI_WARPS + G::N_EXP_WARPS + 1;   // warps (and the MMA lane) that read every job
    static_assert(QT == 2 && !G::REDUX && G::EXP_SPLIT == 1, "the persistent kernel is written for the shipped geometry");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    PSmem<G> &ps = *reinterpret_cast<PSmem<G> *>(smem_raw);
    typename G::Smem &sm = ps.s;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int ws = p.desc_bytes >> 2;   // input word whose byte 3 holds the 8 spare K positions
    const int n_k = ws + 1;             // MMA K steps: 32 K positions = 4 descriptor bytes each

    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            mbar_init(&sm.raw_full[i], 1);
            mbar_init(&sm.b_full[i], G::N_EXP_WARPS);
            mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], 4);
            mbar_init(&ps.a_ready[i], 4);
            mbar_init(&ps.raw_empty[i], G::N_EXP_WARPS);
        }
        for (int i = 0; i < JR; ++i) {
            mbar_init(&ps.sched_full[i], 1);
            mbar_init(&ps.sched_empty[i], N_READERS);
        }
        fence_mbar_init();
    }
    if (COL)
        for (int i = tid; i < QT * 2 * NT; i += THREADS) (&sm.colmin[0][0][0])[i] = KEY_NONE;
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

    // job n as published by the TMA lane; every reader takes a private copy and frees the ring slot at once
    auto read_job = [&](int n, bool whole_warp) {
        const int slot = n % JR;
        mbar_wait_relaxed(&ps.sched_full[slot], (n / JR) & 1, 32);
        const JobInfo ji = ps.jobs[slot];
        if (whole_warp) __syncwarp();
        if (!whole_warp || lane == 0) mbar_arrive(&ps.sched_empty[slot]);
        return ji;
    };
    auto stage_rows = [&](const JobInfo &j, int s) { return min(NT, j.te - (j.tb + s * NT)); };
    auto stage_src = [&](const JobInfo &j, int s) {
        return p.t + (static_cast<size_t>(j.t_row0) + j.tb + s * NT) * p.t_stride;
    };
    auto stage_tma_rows = [&](const JobInfo &j, int s) {
        if (reinterpret_cast<uintptr_t>(stage_src(j, s)) & 15) return 0;
        return stage_rows(j, s) & ~(p.tma_quantum - 1);   // the quantum is a power of two (16 / gcd(stride, 16))
    };
    // accumulator uses of a whole job (two tiles: both once per stage; one tile: the stages alternate)
    auto acc_uses = [&](const JobInfo &j, int acc) {
        return j.n_tiles == 2 ? j.n_stage : (acc == 0 ? (j.n_stage + 1) >> 1 : j.n_stage >> 1);
    };

    if (warp < N_EPI_WARPS) {
        // ============================== epilogue warps ==============================
        // Group g = warps 4g .. 4g+3 owns accumulator g, +-1 tile g and column-minima buffer g for the whole launch and
        // waits for EVERY phase of d_full[g] in order (a warp that sat out some uses of a barrier cannot tell its phases
        // apart by parity).  Two-tile job: group g folds query tile g in every stage.  One-tile job (the last <= 128
        // query rows of a problem): the stages alternate between the accumulators, the last one on accumulator 0, and
        // the groups share the tile's rows, each folding the stages of its own accumulator.
        const int tile = warp >> 2;
        const int quarter = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;  // this warp's TMEM lanes
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t one = static_cast<uint32_t>(p.desc_bytes > 0);  // opaque 1: keeps key adds on the FMA pipe
        const uint32_t scr_addr = smem_u32(sm.scratch[warp]);
        // Look at job n without taking it (the ring slot is freed by read_job at the top of the job's own iteration):
        // when this group builds a +-1 tile for it, start copying the thread's query row into its shared-memory slot
        // (cp.async of the aligned words that hold descriptor bytes: no registers are tied up during the sweep).
        const uint32_t qs_addr = smem_u32(ps.qstage[tid]);
        int q_sh = 0;   // byte misalignment of the staged row
        auto prefetch_query = [&](int n, bool direct) {
            const int slot = n % JR;
            mbar_wait_relaxed(&ps.sched_full[slot], (n / JR) & 1, 32);
            const int4 a = *reinterpret_cast<const int4 *>(&ps.jobs[slot]);            // valid, prob, q_row0, nq
            const int qt0 = ps.jobs[slot].qt0, n_tiles = ps.jobs[slot].n_tiles;
            const bool act = a.x != 0 && tile < n_tiles;
            if (act) {
                const int row_q = qt0 + tile * MQ + row_in_tile;   // rows past the end mirror the last row
                const uintptr_t src =
                    reinterpret_cast<uintptr_t>(p.q + static_cast<size_t>(a.z + min(row_q, a.w - 1)) * p.q_stride);
                q_sh = static_cast<int>(src & 3);
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src & ~static_cast<uintptr_t>(3));
                const int n_need = (q_sh + p.desc_bytes + 3) >> 2;   // aligned words holding descriptor bytes (<= 17)
#pragma unroll
                for (int k = 0; k < 17; ++k)
                    if (k < n_need)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(qs_addr + 4 * k), "l"(s32 + k) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            (void)direct;
            return act;
        };
        // the +-1 tile of the staged query row -> TMEM (see hamming_mma_kernel); the caller has seen every MMA that read
        // the old tile complete
        auto build_a = [&]() {
            uint32_t wq[W];
            {
                asm volatile("cp.async.wait_all;" ::: "memory");
                const int n_need = (q_sh + p.desc_bytes + 3) >> 2;
                const uint32_t *qs = ps.qstage[tid];
                uint32_t lo = qs[0];
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const uint32_t hi = (k + 1 < n_need) ? qs[k + 1] : 0u;
                    wq[k] = __funnelshift_r(lo, hi, q_sh * 8) & word_mask(p.desc_bytes, k);
                    lo = hi;
                }
            }
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
                if (k < n_k) {
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
                    tmem_st8(tmem + lane_base + G::TMEM_A + tile * 128 + 8 * k, a);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ps.a_ready[tile]);
        };

        uint32_t acc_base = 0;   // uses of this group's accumulator before this job = stages the group has folded
        long long pf_wait = 0, pf_build = 0, pf_last = 0, pf_peek = 0;
        const long long pf_estart = PROF_NOW();
        prefetch_query(0, true);
        bool prebuilt = false;
        for (int n = 0;; ++n) {
            const JobInfo cur = read_job(n, true);
            if (!cur.valid) break;
            if (tile < cur.n_tiles && !prebuilt) build_a();   // warp-uniform
            prebuilt = false;
            // the next job's query row, consumed at the end of this job: its latency hides behind the sweep
            const long long tp0 = PROF_NOW();
            const bool nxt_builds = prefetch_query(n + 1, false);
            PROF_ADD(pf_peek, tp0);
            const bool two = cur.n_tiles == 2;
            const int my_n = acc_uses(cur, tile);   // stages this group folds
            {
                const int acc = tile;
                const int first = two ? 0 : ((cur.n_stage - 1 - tile) & 1), step = two ? 1 : 2;
                const int row = cur.qt0 + (two ? tile * MQ : 0) + row_in_tile;
                const uint32_t rbase = static_cast<uint32_t>(cur.qt0 + (two ? tile * MQ : 0) + quarter * 32);
                uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
                for (int i = 0; i < my_n; ++i) {
                    const int s = first + i * step;
                    const uint32_t use = acc_base + i;
                    const int rows = stage_rows(cur, s);
                    const int cm = use & 1;
                    const long long tw = PROF_NOW();
                    mbar_wait_relaxed(&sm.d_full[acc], use & 1, 128);
                    PROF_ADD(pf_wait, tw);
                    tc_fence_after();
                    const long long tl = PROF_NOW();
                    const uint32_t jstage = static_cast<uint32_t>(p.t_index_base + cur.tb + s * NT);
                    uint32_t vv[G::NCH][32];
#pragma unroll
                    for (int h = 0; h < G::NCH; ++h)
                        if (h == 0 || 64 * h < rows) tmem_ld64_packed(tmem + lane_base + acc * NT + 64 * h, vv[h]);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.d_empty[acc]);
                    if (i == my_n - 1 && nxt_builds) {
                        // Two tiles: d_full of the job's last stage — every MMA that read this group's +-1 tile is
                        // complete.  One tile: group 0's last stage is the job's last (its commit covers all earlier
                        // MMAs), and tile 1 is not read by this job at all.
                        const long long tb0 = PROF_NOW();
                        build_a();
                        PROF_ADD(pf_build, tb0);
                        PROF_ADD(pf_last, tl);
                        prebuilt = true;
                    }
#pragma unroll
                    for (int h = 0; h < G::NCH; ++h) {
                        if (h > 0 && 64 * h >= rows) break;  // warp-uniform
                        uint32_t (&v)[32] = vv[h];
                        if (COL) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t a = scr_addr + lane * 128 + ((i ^ (lane & 7)) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[4 * i]), "r"(v[4 * i + 1]),
                                             "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                                             : "memory");
                            }
                        }
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
                                top2_insert(c[i] >= (505u << 7) ? KEY_NONE : ((c[i] >> 7) << KEY_IDX_BITS) + jh + (c[i] & 127u),
                                            b1, b2);
                        }
                        if (COL) {
                            uint32_t m = 0xFFFFFFFFu;
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < 32; r += 2) {
                                uint32_t x0, x1;
                                const uint32_t a0 = scr_addr + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2));
                                const uint32_t a1 =
                                    scr_addr + (r + 1) * 128 + ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2));
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
                                if (c < rows)
                                    atomicMin(&sm.colmin[tile][cm][c], ((k16 >> 5) << KEY_IDX_BITS) + rbase + (k16 & 31u));
                            }
                        }
                    }
                    if (COL) {
                        // the 4 warps of this query tile merge: one global atomicMin per train row, tile and stage
                        if (tile == 0)
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                        else
                            asm volatile("bar.sync 2, 128;" ::: "memory");
                        for (int c = row_in_tile; c < rows; c += 128) {
                            atomicMin(p.col_keys + cur.t_row0 + cur.tb + s * NT + c, sm.colmin[tile][cm][c]);
                            sm.colmin[tile][cm][c] = KEY_NONE;  // reused two stages on, after the next stage's barrier
                        }
                    }
                }
                if (my_n > 0 && row < cur.nq) {
                    const size_t orow = static_cast<size_t>(cur.out_row0) + row;
                    const bool sole = p.jobs_y == 1 && two;   // nobody else holds results for this row
                    if (!TOP2 && p.compact) {
                        uint32_t *c = reinterpret_cast<uint32_t *>(p.row_keys) + orow;
                        if (sole)
                            *c = b1;
                        else
                            atomicMin(c, b1);
                    } else {
                        uint2 *g = p.row_keys + orow;
                        if (sole)
                            *g = make_uint2(b1, b2);
                        else if (TOP2)
                            merge_row_keys(g, b1, b2);
                        else
                            atomicMin(&g->x, b1);  // second key stays KEY_NONE (pre-set)
                    }
                }
            }
            acc_base += my_n;
        }
#ifdef SLAMFE_MMA_PROF
        if (blockIdx.x < 256 && lane == 0 && quarter == 0) {
            unsigned long long *o = g_mma_prof[blockIdx.x] + 9 + 3 * tile;   // 9..11 group 0, 12..14 group 1
            o[0] = pf_wait; o[1] = pf_build; o[2] = pf_last;
            if (tile == 0) g_mma_prof[blockIdx.x][15] = pf_peek;
        }
        (void)pf_estart;
#endif
    } else if (warp < MMA_WARP) {
        // ============================== expander warps ==============================
        const int r = tid - N_EPI_WARPS * 32;                // this thread's train row within every stage
        uint32_t mul[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) mul[s] = (1u << (7 - s)) * static_cast<uint32_t>(p.desc_bytes > 0);
        const uint32_t last_mask = word_mask(p.desc_bytes, ws);   // the word that holds the descriptor's tail + the spare byte
        auto sweep = [&](auto rows64_c) {
            constexpr bool ROWS64 = decltype(rows64_c)::value;
            uint32_t gsb = 0;                 // B stages before this job
            static_assert(NB == 2, "two raw buffers");
            uint32_t raw_cnt0 = 0, raw_cnt1 = 0;  // bulk copies seen per raw buffer (a stage the copy cannot take arms nothing)
            for (int n = 0;; ++n) {
                const JobInfo cur = read_job(n, true);
                if (!cur.valid) break;
                for (int s = 0; s < cur.n_stage; ++s) {
                    const uint32_t g = gsb + s;
                    const int b = g % NB;
                    const uint32_t ph = (g / NB) & 1;
                    const int rows = stage_rows(cur, s), trows = stage_tma_rows(cur, s);
                    if (trows > 0) {
                        mbar_wait(&sm.raw_full[b], (b ? raw_cnt1 : raw_cnt0) & 1);
                        if (b) ++raw_cnt1; else ++raw_cnt0;
                    }
                    mbar_wait(&sm.b_empty[b], ph ^ 1);
                    // 64-byte rows (see hamming_mma_kernel): conflict-free 16-byte chunk reads, words straight to their K positions
                    const bool fast64 = ROWS64 && r < trows && r < rows;
                    if (fast64) {
                        uint8_t *dst64 = sm.b[b] + r * 16;
                        const uint4 *raw128 = reinterpret_cast<const uint4 *>(sm.raw[b] + r * 64);
                        const uint32_t slo = 0x80000000u;
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
                                const uint32_t hi = raw32[k + 1];
                                w[k] = __funnelshift_r(lo, hi, sh);
                                lo = hi;
                            }
                        } else {  // rows the bulk copy could not take (misaligned source / tail): read global memory
                            load_desc_global(stage_src(cur, s) + static_cast<size_t>(r) * p.t_stride, p.desc_bytes, w);
                        }
                        uint8_t *dst = sm.b[b] + r * 16;
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
                    fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();        // ... of every lane, before the warp's one arrival
                    if (lane == 0) {
                        if (trows > 0) mbar_arrive(&ps.raw_empty[b]);   // the staging buffer may take its next bulk copy
                        mbar_arrive(&sm.b_full[b]);
                    }
                }
                gsb += cur.n_stage;
            }
        };
        if (n_k == W && p.t_stride == 64)
            sweep(std::true_type{});
        else
            sweep(std::false_type{});
    } else if (warp == MMA_WARP) {
        // ============================== MMA issue (one elected lane) ===========
        if (elect_one()) {
            const uint64_t desc0 = umma_desc(smem_u32(sm.b[0]), LBO, 128);
            uint32_t gsb = 0, acc_base0 = 0, acc_base1 = 0, a_cnt0 = 0, a_cnt1 = 0;
            // everything MMA job (s, t) of CTA job j needs: the +-1 tile (first stage), its B stage (first tile only),
            // its accumulator
            long long pf_a = 0, pf_b = 0, pf_d = 0, pf_read = 0, pf_mid = 0, pf_jobs = 0, pf_mmas = 0, pf_a1 = 0;
            auto wait_mma = [&](const JobInfo &j, int s, int t) {
                long long t0 = PROF_NOW();
                if (s == 0) {
                    mbar_wait(&ps.a_ready[t], (t ? a_cnt1 : a_cnt0) & 1);
                    if (t) ++a_cnt1; else ++a_cnt0;
                    if (t) PROF_ADD(pf_a1, t0); else PROF_ADD(pf_a, t0);
                }
                t0 = PROF_NOW();
                const uint32_t g = gsb + s;
                if (t == 0) mbar_wait(&sm.b_full[g % NB], (g / NB) & 1);
                PROF_ADD(pf_b, t0);
                t0 = PROF_NOW();
                const int acc = j.n_tiles == 2 ? t : ((j.n_stage - 1 - s) & 1);   // one tile: the last stage on accumulator 0
                const uint32_t use = (acc ? acc_base1 : acc_base0) + (j.n_tiles == 2 ? s : (s >> 1));
                mbar_wait(&sm.d_empty[acc], (use & 1) ^ 1);
                PROF_ADD(pf_d, t0);
                tc_fence_after();
            };
            int n = 0, s = 0, t = 0;
            JobInfo cur = read_job(0, false);
            const long long pf_start = PROF_NOW();
            if (cur.valid) wait_mma(cur, 0, 0);
            while (cur.valid) {
                const uint32_t g = gsb + s;
                const int b = g % NB, acc = cur.n_tiles == 2 ? t : ((cur.n_stage - 1 - s) & 1);
                const uint64_t desc_b = desc0 + static_cast<uint64_t>(b * (G::B_STAGE >> 4));
                const uint32_t d_addr = tmem + acc * NT, a_addr = tmem + G::TMEM_A + t * 128;
                int ns = s, nt = t + 1;
                if (nt == cur.n_tiles) {
                    nt = 0;
                    ns = s + 1;
                }
                const bool last = ns == cur.n_stage;   // the job's last MMA job: the next +-1 tile is built only after
                                                       // these MMAs complete, so its barrier is waited for afterwards
                if (n_k == 16) {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    {
                        const long long t0 = PROF_NOW();
                        if (!last) wait_mma(cur, ns, nt);
                        PROF_ADD(pf_mid, t0);
                    }
                    ++pf_mmas;
#pragma unroll
                    for (int k = 12; k < 16; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, 1);
                } else {
                    for (int k = 0; k < n_k; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    if (!last) wait_mma(cur, ns, nt);
                }
                if (t == cur.n_tiles - 1) umma_commit(&sm.b_empty[b]);
                umma_commit(&sm.d_full[acc]);
                if (last) {
                    gsb += cur.n_stage;
                    acc_base0 += acc_uses(cur, 0);
                    acc_base1 += acc_uses(cur, 1);
                    const long long t0 = PROF_NOW();
                    cur = read_job(++n, false);
                    PROF_ADD(pf_read, t0);
                    ++pf_jobs;
                    s = 0;
                    t = 0;
                    if (cur.valid) wait_mma(cur, 0, 0);
                } else {
                    s = ns;
                    t = nt;
                }
            }
#ifdef SLAMFE_MMA_PROF
            if (blockIdx.x < 256) {
                unsigned long long *o = g_mma_prof[blockIdx.x];
                o[0] = clock64() - pf_start; o[1] = pf_read; o[2] = pf_a; o[3] = pf_b; o[4] = pf_d; o[5] = pf_mid;
                o[6] = pf_jobs; o[7] = pf_mmas; o[8] = pf_a1;
            }
#endif
        }
    } else {
        // ============================== job fetch + TMA producer (one elected lane) ===========
        if (elect_one()) {
            bool exhausted = false;
            auto fetch = [&](int n) {
                JobInfo ji{};
                while (!exhausted) {
                    const uint32_t id = atomicAdd(p.job_counter, 1u);
                    if (id >= static_cast<uint32_t>(p.jobs_total)) {
                        exhausted = true;
                        // every CTA fetches exactly once past the end: this one is the launch's last fetch
                        if (id == static_cast<uint32_t>(p.jobs_total) + gridDim.x - 1) atomicExch(p.job_counter, 0u);
                        break;
                    }
                    const int x = static_cast<int>(id % static_cast<uint32_t>(p.jobs_x));
                    const int yz = static_cast<int>(id / static_cast<uint32_t>(p.jobs_x));
                    const int y = yz % p.jobs_y, prob = yz / p.jobs_y;
                    int q_row0 = 0, nq = p.nq, t_row0 = 0, nt = p.nt;
                    if (p.q_off) {
                        q_row0 = p.q_off[prob];
                        nq = p.q_cnt ? p.q_cnt[prob] : p.q_off[prob + 1] - q_row0;
                    }
                    if (p.t_off) {
                        t_row0 = p.t_off[prob];
                        nt = p.t_cnt ? p.t_cnt[prob] : p.t_off[prob + 1] - t_row0;
                    }
                    const int qt0 = x * CQ, tb = y * p.t_slice;
                    if (qt0 >= nq || tb >= nt) continue;   // an empty tile of the ragged job space
                    ji.valid = 1; ji.prob = prob; ji.q_row0 = q_row0; ji.nq = nq; ji.qt0 = qt0; ji.t_row0 = t_row0;
                    ji.tb = tb; ji.te = min(nt, tb + p.t_slice);
                    ji.n_stage = (ji.te - tb + NT - 1) / NT;
                    ji.n_tiles = nq - qt0 > MQ ? 2 : 1;
                    ji.out_row0 = p.row_out_off ? p.row_out_off[prob] : q_row0;
                    break;
                }
                const int slot = n % JR;
                if (n >= JR) mbar_wait_relaxed(&ps.sched_empty[slot], ((n / JR) - 1) & 1, 64);
                ps.jobs[slot] = ji;
                mbar_arrive(&ps.sched_full[slot]);
                return ji;
            };
            uint32_t gsb = 0;
            uint32_t armed0 = 0, armed1 = 0;   // bulk copies issued per staging buffer
            JobInfo cur = fetch(0);
            for (int n = 0; cur.valid; ++n) {
                JobInfo nxt{};
                const int fetch_at = min(cur.n_stage, NB) - 1;   // the next job is fetched once this one's first stages are on their way
                for (int s = 0; s < cur.n_stage; ++s) {
                    const int b = (gsb + s) % NB;
                    const int trows = stage_tma_rows(cur, s);
                    if (trows > 0) {
                        // raw[b] is free again once every expander has read its previous copy.  raw_empty counts bulk
                        // copies only (stages whose source the bulk copy cannot take arm nothing and the expanders run
                        // through them without this lane), so its phases are waited for one by one, never skipped.
                        const uint32_t armed = b ? armed1 : armed0;
                        if (armed > 0) mbar_wait_relaxed(&ps.raw_empty[b], (armed - 1) & 1, 256);
                        if (b) ++armed1; else ++armed0;
                        const uint32_t bytes = static_cast<uint32_t>(trows) * p.t_stride;
                        mbar_arrive_expect_tx(&sm.raw_full[b], bytes);
                        tma_load_1d(sm.raw[b], stage_src(cur, s), bytes, &sm.raw_full[b]);
                    }
                    if (s == fetch_at) nxt = fetch(n + 1);
                }
                gsb += cur.n_stage;
                cur = nxt;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <class G, bool COL, bool TOP2>
int launch_mma_persistent(const HammingParams &p, int ctas, cudaStream_t stream)
{
    static bool configured = false;  // per instantiation; racing threads set the same value
    constexpr int smem = static_cast<int>(sizeof(PSmem<G>));
    static_assert(smem <= 227 * 1024, "shared memory per CTA");
    if (!configured) {
        SLAMFE_CUDA_OK(cudaFuncSetAttribute(hamming_mma_persistent_kernel<G, COL, TOP2>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    hamming_mma_persistent_kernel<G, COL, TOP2><<<ctas, G::THREADS, smem, stream>>>(p);
    return launch_status();
}

// The job counter of one persistent launch.  A counter is zero whenever no launch uses it (the kernel's last fetch
// resets it), so the library hands out the words of a per-device ring in turn: launches on different streams never
// share a counter unless 4096 launches lie between them.
uint32_t *next_job_counter()
{
    constexpr int RING = 4096, MAX_DEV = 64;
    static uint32_t *ring[MAX_DEV] = {};
    static unsigned next[MAX_DEV] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!ring[dev]) {
        uint32_t *mem = nullptr;
        if (cudaMalloc(&mem, RING * sizeof(uint32_t)) != cudaSuccess) return nullptr;
        if (cudaMemset(mem, 0, RING * sizeof(uint32_t)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            cudaFree(mem);
            return nullptr;
        }
        ring[dev] = mem;
    }
    return ring[dev] + (next[dev]++ % RING);
}

