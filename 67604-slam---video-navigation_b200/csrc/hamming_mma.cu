// Brute-force Hamming matcher on the 5th-generation tensor cores (tcgen05, sm_100a) — same contract,
// same keys, same tie-breaks as hamming_top2_kernel in hamming.cu (reference call sites:
// final_project/backend/database/database.py:54-55, backend/loop/loop_closure.py:422,
// final_project/algorithms/matching.py:15,44, VAN_ex/code/ex1.py:189-190).
//
// Hamming distance as an exact integer contraction.  With query bits q_k and train bits t_k,
//     d(q, t) = sum_k q_k + sum_k t_k * (1 - 2 q_k) = popc(q) + A_q . B_t,
// A_q = (1 - 2 q_k) in {+1, -1} (int8), B_t = t_k.  tcgen05.mma kind::i8 accumulates A . B^T exactly
// in int32, so distance, key = (d << 22) | index and every tie-break are bit-identical to the
// XOR/POPC kernel.  Three refinements make the accumulator directly usable as a 16-bit key field:
//   * B_t = 128 * t_k (uint8 0x80): the bit -> byte expansion is one IMAD (w << (7 - s), FMA pipe) and
//     one LOP3 (& 0x80808080, ALU pipe) per four K positions, and the accumulator is 128 * (...), i.e.
//     already shifted past a 7-bit index field;
//   * popc(q) rides in spare K positions: descriptor byte 4*ws + 3 (ws = desc_bytes / 4) lies beyond
//     the descriptor, its 8 K positions are free.  Four of them carry A = c_i (sum c_i = popc(q),
//     c_i <= 127) against B = 0x80 on every valid train row, so acc = 128 * d >= 0 and d needs no
//     per-pair fix-up;
//   * the other four carry A = 127 against B = 0x80 on INVALID train rows only (rows past the end of
//     the last stage, whose data bits are zero): acc = 128 * 508, larger than any real distance
//     (d <= 8 * 63), so ragged stages need no masking in the epilogue.
// The order of the K positions is irrelevant as long as A and B agree: output word s of input word w
// holds bits s, s+8, s+16, s+24 of w.  Needs desc_bytes <= 63 (AKAZE: 61); 64-byte descriptors take
// the INT kernel.
//
// One CTA = 256 query rows (two UMMA M = 128 tiles, both +-1 tiles resident in TMEM as the A operands)
// against a train sweep in stages of 128 rows (UMMA N = 128):
//   warps 0-7   epilogue, warp w owns TMEM lane quarter w & 3 of query tile w >> 2.  Once: write the
//               query tile into TMEM (A operand, 128 columns per tile).  Per stage and 64 columns: one
//               tcgen05.ld (.pack::16b, 64 columns in 32 registers), key16 = acc + column (one IMAD per two
//               pairs), running minimum with VIMNMX3.U16x2 (one per four pairs), folded into the 32-bit row
//               keys once per stage.  Column minima (crossCheck / backward match): the warp transposes its
//               32 x 64 block of distances through a swizzled shared-memory scratch, every lane reduces
//               two train columns over the warp's 32 query rows (key16 = acc/4 + row), the 4 warps of a
//               tile merge through shared-memory atomics, one global atomicMin per train row, tile and stage;
//   warps 8-11  expanders: raw 61-byte train rows (1-D TMA bulk copy, double buffered) -> 0x80/0 bytes in
//               the K-major SWIZZLE_NONE core-matrix layout of the B operand (8 rows x 16 B contiguous,
//               K chunks LBO = 128*16 B apart), one row per thread, conflict-free STS.128;
//   warp 12     issues the TMA copies and 2 x 16 tcgen05.mma (M128 N128 K32, A from TMEM) per stage from one
//               elected lane; tcgen05.commit hands the B stage back to the expanders and the accumulator
//               to the epilogue.  The two query tiles' accumulators (2 x 128 TMEM columns) double-buffer
//               each other: while the epilogue drains tile 0 the tensor core works on tile 1 of the same
//               B stage (a CTA with a single tile alternates between the two accumulators instead).
// Expanding a train row once per 256 query rows (v1: per 128) and the packed epilogue take the INT work
// around the tensor pipe from ~6000 to ~1000 warp instructions per 16 K pairs (profiles/r02_*).  The MMA
// issue loop is fully unrolled with precomputed descriptors: a tcgen05.mma that takes 64 cycles leaves the
// issuing thread no room for per-instruction descriptor arithmetic (scripts/probe_tcgen05.cu rate2).
// Facts pinned on a B200 by scripts/probe_tcgen05.cu (profiles/r02_probe_tcgen05.log): descriptor
// field meaning (LBO = K-chunk stride, SBO = 8-row stride), TMEM A layout (4 K-bytes per column),
// exactness, MMA cycles at the issue floor, TMEM read bandwidth.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"

#include "hamming_params.cuh"

namespace slamfe {

namespace {

constexpr int MQ = 128;                 // query rows per UMMA tile (M)
constexpr int KCH = 32;                 // 16-byte K chunks per row (512 K positions)
constexpr uint32_t TMEM_COLS = 512;

// Geometry of one instantiation.  QT query tiles of 128 rows per CTA (their +-1 tiles live in TMEM as A
// operands), train stages of NT rows (= UMMA N), two accumulators of NT columns:
//   <2, 128>  256 query rows per CTA: a train row is expanded once per 256 queries; the two tiles'
//             accumulators double-buffer each other; column minima by shared-memory transpose or redux
//   <1, 192>  128 query rows per CTA, N = 192: fewer, larger MMAs (a tcgen05.mma with A in TMEM costs
//             ~83 + 0.2 N cycles on this part: 109 at N = 128, 122 at N = 192) at twice the expansion work
//             per pair; the accumulators alternate between stages; column minima by redux only (no room
//             for the transpose scratch next to 2 x 96 KB of B stages)
template <int QT_, int NT_, bool REDUX_, int EXP_SPLIT_ = 1, bool SCR2_ = false, bool LD16_ = false, bool IDXK_ = false,
          int ESPLIT_ = 1>
struct Geo {
    static constexpr int ESPLIT = ESPLIT_;           // epilogue warps per TMEM lane quarter and query tile: 2 = each takes one
                                                     // 64-column chunk of the accumulator (16 epilogue warps, 80 registers)
    static constexpr bool IDXK = IDXK_;              // the train row's index within the stage rides in a spare K position
                                                     // (A = 1, B = index): the accumulator IS the 16-bit row key, no add
    static constexpr bool LD16 = LD16_;              // epilogue reads the accumulators with tcgen05.ld.16x256b: a thread holds
                                                     // 4 query rows x 32 train columns and pre-reduces rows in registers
    static constexpr bool SCR2 = SCR2_;              // one transpose scratch per 64-column chunk: both chunks of a stage in flight
    static constexpr int QT = QT_, NT = NT_;
    static constexpr bool REDUX = REDUX_;
    static constexpr int CQ = MQ * QT;               // query rows per CTA
    static constexpr int LBO = NT * 16;              // bytes between consecutive K chunks of the B tile
    static constexpr int B_STAGE = KCH * LBO;
    static constexpr int NB = 2;                     // B (and raw) stages
    static constexpr int RAW_STAGE = NT * SLAMFE_MAX_DESC_BYTES + 16;
    static constexpr int EXP_SPLIT = EXP_SPLIT_;       // expander threads per train row (2 measured slower than 1:
                                                       // the expansion is bound by shared pipes, not by latency)
    static constexpr int N_EPI_WARPS = 4 * QT * ESPLIT, N_EXP_WARPS = NT / 32 * EXP_SPLIT;
    static constexpr int MMA_WARP = N_EPI_WARPS + N_EXP_WARPS;
    static constexpr int TMA_WARP = MMA_WARP + 1;
    static constexpr int THREADS = (TMA_WARP + 1) * 32;
    static constexpr int NCH = NT / 64;              // 64-column chunks of an accumulator
    static constexpr uint32_t TMEM_A = 2 * NT;       // first column of the query tiles
    static constexpr uint32_t IDESC = (2u << 4)                                 // D format S32
                                      | (1u << 7)                               // A signed 8-bit
                                      | (0u << 10)                              // B unsigned 8-bit
                                      | (static_cast<uint32_t>(NT >> 3) << 17)  // N
                                      | (static_cast<uint32_t>(MQ >> 4) << 24); // M; both operands K-major
    static_assert(2 * NT + 128 * QT <= 512, "TMEM: two accumulators + the query tiles");
    static_assert(NT % 64 == 0 && NT % 32 == 0, "stage = whole 64-column chunks");
    struct __align__(16) Smem {
        uint8_t b[NB][B_STAGE];
        uint8_t raw[NB][RAW_STAGE];
        uint32_t scratch[REDUX ? 1 : N_EPI_WARPS][REDUX ? 4 : (SCR2 ? 2 : 1) * 32 * 32];  // per-warp 32 x 64 packed distances, swizzled
        uint32_t colmin[QT][2][NT];                                      // per query tile, double buffered over stages
        uint64_t raw_full[NB], b_full[NB], b_empty[NB], d_full[2], d_empty[2], a_ready;
        uint32_t tmem_base;
    };
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait with back-off for waiters that have slack (epilogue on d_full, TMA producer): a plain try_wait spin
// loop competes for issue slots with the expander warp on the same scheduler (measured: the expansion of a
// stage took 2350 cycles with spinning neighbours, the MMAs it feeds 2048).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, unsigned ns)
{
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
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
// 64 accumulator columns of this warp's 32 lanes as 32 registers: register j = column 2j (low half) and
// column 2j + 1 (high half); the accumulators are < 2^16 by construction.
#ifndef SLAMFE_MMA_PACK16
#define SLAMFE_MMA_PACK16 1
#endif
#define SLAMFE_LD32_OPERANDS                                                                                          \
    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"      \
    "%29,%30,%31}, [%32];"                                                                                            \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
        : "r"(taddr)                                                                                                  \
        : "memory"
__device__ __forceinline__ void tmem_ld64_packed(uint32_t taddr, uint32_t (&v)[32])
{
#if SLAMFE_MMA_PACK16
    // .pack::16b: two adjacent 32-bit columns -> one register (low 16 bits of each), done by the load path
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " SLAMFE_LD32_OPERANDS);
#else
    // plain loads, packed with one IMAD per column pair
    uint32_t out[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " SLAMFE_LD32_OPERANDS);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) out[16 * half + j] = v[2 * j + 1] * 65536u + v[2 * j];
        taddr += 32;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = out[j];
#endif
}
// 16 TMEM lanes x 128 columns as 32 packed registers (profiles/r02_probe_ldshape.log): with q = lane / 4, m = lane % 4,
// register 4u + k holds columns 16u + 4m + 2(k & 1) (low half) and + 1 (high half) of TMEM lane q + 8(k >> 1)
__device__ __forceinline__ void tmem_ld16x128_packed(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.pack::16b.b32 " SLAMFE_LD32_OPERANDS);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// bits s, s+8, s+16, s+24 of w as four 0x80 / 0 bytes; `mul` = 1 << (7 - s) lives in a register so that the shift is an
// IMAD on the FMA pipe (ptxas would turn a constant power of two into an ALU shift)
__device__ __forceinline__ uint32_t spread80(uint32_t w, uint32_t mul)
{
    uint32_t x;
    asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(x) : "r"(w), "r"(mul));
    return x & 0x80808080u;
}
__device__ __forceinline__ uint32_t add_imad(uint32_t x, uint32_t one, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(c));
    return r;
}

// merge a candidate key into a sorted (b1 <= b2) pair
__device__ __forceinline__ void top2_insert(uint32_t key, uint32_t &b1, uint32_t &b2)
{
    b2 = min(b2, max(b1, key));
    b1 = min(b1, key);
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// accumulator used by (stage s, query tile t) and how many times it was used before
__device__ __forceinline__ int acc_of(int s, int t, int n_tiles) { return n_tiles == 2 ? t : (s & 1); }
__device__ __forceinline__ int acc_use(int s, int n_tiles) { return n_tiles == 2 ? s : (s >> 1); }

// grid = (query tiles of G::CQ rows, train slices, problems)
template <class G, bool COL, bool TOP2>
__global__ void __launch_bounds__(G::THREADS, 1) hamming_mma_kernel(const HammingParams p)
{
    constexpr int QT = G::QT, NT = G::NT, CQ = G::CQ, LBO = G::LBO, NB = G::NB;
    constexpr int N_EPI_WARPS = G::N_EPI_WARPS, MMA_WARP = G::MMA_WARP, THREADS = G::THREADS;
    using Smem = typename G::Smem;
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
    const int qt0 = blockIdx.x * CQ;
    const int tb = blockIdx.y * p.t_slice;
    if (qt0 >= nq || tb >= nt) return;  // CTA-uniform; outputs were pre-set to KEY_NONE
    const int te = min(nt, tb + p.t_slice);
    const int n_stage = (te - tb + NT - 1) / NT;
    const int ws = p.desc_bytes >> 2;   // input word whose byte 3 holds the 8 spare K positions
    const int n_k = ws + 1;             // MMA K steps: 32 K positions = 4 descriptor bytes each
    const int n_tiles = (QT == 2 && nq - qt0 > MQ) ? 2 : 1;

    // epilogue threads: start fetching their query row now, its HBM latency overlaps the TMEM allocation and
    // the barrier set-up below
    uint32_t wq[W];
    const bool is_epi = warp < N_EPI_WARPS && ((warp >> 2) % QT) < n_tiles;
    if (is_epi) {
        const int row_q = qt0 + ((warp >> 2) % QT) * MQ + (warp & 3) * 32 + lane;
        load_desc_global(p.q + static_cast<size_t>(q_row0 + min(row_q, nq - 1)) * p.q_stride, p.desc_bytes, wq);
    }
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            mbar_init(&sm.raw_full[i], 1);
            mbar_init(&sm.b_full[i], G::N_EXP_WARPS * 32);
            mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], 128 * G::ESPLIT);
        }
        mbar_init(&sm.a_ready, n_tiles * 128 * G::ESPLIT);
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

    const uint8_t *t_base = p.t + static_cast<size_t>(t_row0) * p.t_stride;
    auto stage_rows = [&](int s) { return min(NT, te - (tb + s * NT)); };
    auto stage_src = [&](int s) { return t_base + static_cast<size_t>(tb + s * NT) * p.t_stride; };
    auto stage_tma_rows = [&](int s) {
        if (reinterpret_cast<uintptr_t>(stage_src(s)) & 15) return 0;
        const int rows = stage_rows(s);
        return rows & ~(p.tma_quantum - 1);   // the quantum is a power of two (16 / gcd(stride, 16))
    };

    if (warp < N_EPI_WARPS) {
        // ============================== epilogue warps ==============================
        const int tile = (warp >> 2) % QT;
        const int chunk = (warp >> 2) / QT;   // ESPLIT == 2: this warp's 64-column chunk of every accumulator (and its K half
                                              // of the +-1 tile); ESPLIT == 1: 0, the warp takes both chunks
        if (tile < n_tiles) {  // warp-uniform: a CTA over the last <= 128 query rows has a single tile
            const int quarter = warp & 3;
            const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;  // this warp's TMEM lanes
            const int row_in_tile = quarter * 32 + lane;
            const int row = qt0 + tile * MQ + row_in_tile;
            // rows past the end mirror the last row (min(row, nq - 1) at kernel entry): their results are not
            // written, and for column minima they lose every tie to the real row
            const uint32_t one = static_cast<uint32_t>(p.desc_bytes > 0);  // opaque 1: keeps key adds on the FMA pipe
            // column-minima key of a packed accumulator pair: (d << 5) | query row within the warp.  With the train index in
            // the accumulator's low 7 bits (IDXK) the row field is masked in with one LOP3 instead of added with one IMAD
            auto col_key = [&](uint32_t acc2, uint32_t rowc) {
                if (G::IDXK) {
                    uint32_t r;
                    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(acc2 >> 2), "r"(0x3FE03FE0u), "r"(rowc));   // (a & b) | c: the mask also drops the two index bits
                                                                                                                     // the shift moved from the high half into the low one
                    return r;
                }
                return add_imad(acc2 >> 2, one, rowc);
            };
            {
                uint32_t (&w)[W] = wq;   // fetched at kernel entry
                uint32_t pq = 0;
#pragma unroll
                for (int k = 0; k < W; ++k) pq += __popc(w[k]);
                // popc(q) split over the four spare positions s = 0..3; 127 on the invalid-row markers s = 4..7
                uint32_t spare[8];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint32_t c = min(pq, 127u);
                    spare[s] = c << 24;
                    pq -= c;
                }
#pragma unroll
                for (int s = 4; s < 8; ++s) spare[s] = 127u << 24;
                if (G::IDXK) {   // invalid rows: 2 x 127 x 255 = 128 * 506 + 2; position 7: + index of the train row
                    spare[6] = 0u;
                    spare[7] = 1u << 24;
                }
                // +-1 bytes: bit s of every byte moves to the byte's msb (IMAD by 1 << (7 - s), FMA pipe), PRMT
                // replicates it over the byte (0xFF / 0x00), OR 1 makes it -1 / +1: 2 ALU-pipe ops per 4 positions
                uint32_t mul[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) mul[s] = (1u << (7 - s)) * one;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    if (k < n_k && (G::ESPLIT == 1 || (k >> 3) == chunk)) {
                        uint32_t a[8];
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            uint32_t t, m;
                            asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(t) : "r"(w[k]), "r"(mul[s]));
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
                mbar_arrive(&sm.a_ready);
            }
            uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
            uint32_t rb1[4] = {KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE}, rb2[4] = {KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE};  // LD16
            const uint32_t scr_addr = smem_u32(sm.scratch[G::REDUX ? 0 : warp]);
            const uint32_t rbase = static_cast<uint32_t>(qt0 + tile * MQ + quarter * 32);
            for (int s = 0; s < n_stage; ++s) {
                const int acc = acc_of(s, tile, n_tiles), use = acc_use(s, n_tiles);
                const int rows = stage_rows(s);
                mbar_wait_relaxed(&sm.d_full[acc], use & 1, 128);
                tc_fence_after();
                const uint32_t jstage = static_cast<uint32_t>(p.t_index_base + tb + s * NT);
                if constexpr (G::LD16) {
                    static_assert(!G::LD16 || NT == 128, "16x256b.x8.pack covers 128 columns");
                    // Thread (q, m) = (lane / 4, lane % 4) holds query rows q, q + 8, q + 16, q + 24 of the warp's 32 and
                    // train columns 16u + 4m + {0..3}, u = 0..7.  Row minima stay in the thread (four running keys,
                    // merged over m at the end of the sweep); column minima are reduced over the thread's four rows
                    // in registers BEFORE anything goes through shared memory: 2 KB out + 2 KB back per warp and stage
                    // instead of 8 + 8 KB.
                    const int q = lane >> 2, m4 = lane & 3;
                    uint32_t va[32], vb[32];
                    tmem_ld16x128_packed(tmem + lane_base + acc * NT, va);
                    tmem_ld16x128_packed(tmem + lane_base + (16u << 16) + acc * NT, vb);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(&sm.d_empty[acc]);
                    const uint32_t mcol = static_cast<uint32_t>(4 * m4) * 0x00010001u;
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri) {
                        const uint32_t (&src)[32] = ri < 2 ? va : vb;
                        const int rs = (ri & 1) * 2;
                        if (!TOP2) {
                            uint32_t m = 0xFFFFFFFFu;
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const uint32_t k0 = add_imad(src[4 * u + rs], one, ((16u * u + 1u) << 16) | (16u * u));
                                const uint32_t k1 = add_imad(src[4 * u + rs + 1], one, ((16u * u + 3u) << 16) | (16u * u + 2u));
                                m = __vimin3_u16x2(m, k0, k1);
                            }
                            m += mcol;
                            const uint32_t k16 = min(m & 0xFFFFu, m >> 16);
                            rb1[ri] = min(rb1[ri], ((k16 >> 7) << KEY_IDX_BITS) + jstage + (k16 & 127u));
                        } else {
                            uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;  // per 16-bit half: best and second best
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk) {
                                    const uint32_t k = add_imad(src[4 * u + rs + kk], one,
                                                                ((16u * u + 2u * kk + 1u) << 16) | (16u * u + 2u * kk));
                                    m2 = __vminu2(m2, __vmaxu2(m1, k));
                                    m1 = __vminu2(m1, k);
                                }
                            }
                            const uint32_t c[4] = {m1 & 0xFFFFu, m1 >> 16, m2 & 0xFFFFu, m2 >> 16};
#pragma unroll
                            for (int i = 0; i < 4; ++i)  // empty slots and invalid train rows (d = 508) are no candidates
                                top2_insert(c[i] >= (505u << 7) ? KEY_NONE
                                                                : ((c[i] >> 7) << KEY_IDX_BITS) + jstage + (c[i] & 127u) + 4u * m4,
                                            rb1[ri], rb2[ri]);
                        }
                    }
                    if (COL) {
                        // key16 = (d << 5) | row-in-warp; the thread's four rows are q + 8 ri
                        const uint32_t rq = static_cast<uint32_t>(q) * 0x00010001u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {   // four registers (two units) per 16-byte store
                            uint32_t cmn[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int u = 2 * i + (jj >> 1), kk = jj & 1;
                                const uint32_t x0 = add_imad(va[4 * u + kk] >> 2, one, rq);
                                const uint32_t x1 = add_imad(va[4 * u + 2 + kk] >> 2, one, rq + 0x00080008u);
                                const uint32_t x2 = add_imad(vb[4 * u + kk] >> 2, one, rq + 0x00100010u);
                                const uint32_t x3 = add_imad(vb[4 * u + 2 + kk] >> 2, one, rq + 0x00180018u);
                                cmn[jj] = __vminu2(__vimin3_u16x2(x0, x1, x2), x3);
                            }
                            // slot of thread t: 64 bytes, 16-byte chunk i stored at chunk i ^ ((t >> 1) & 3): conflict-free STS.128
                            const uint32_t a = scr_addr + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(cmn[0]), "r"(cmn[1]), "r"(cmn[2]),
                                         "r"(cmn[3])
                                         : "memory");
                        }
                        __syncwarp();
                        // thread t = (q, m) merges registers 2q, 2q + 1 of the eight threads (q', m): train columns 4t .. 4t + 3
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const int j = 2 * q + jj;
                            uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
                            for (int qq = 0; qq < 8; qq += 2) {
                                uint32_t x0, x1;
                                const int t0 = 4 * qq + m4, t1 = 4 * (qq + 1) + m4;
                                const uint32_t a0 = scr_addr + t0 * 64 + ((((j >> 2) ^ ((t0 >> 1) & 3)) << 4) | ((j & 3) << 2));
                                const uint32_t a1 = scr_addr + t1 * 64 + ((((j >> 2) ^ ((t1 >> 1) & 3)) << 4) | ((j & 3) << 2));
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(a0) : "memory");
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x1) : "r"(a1) : "memory");
                                mm = __vimin3_u16x2(mm, x0, x1);
                            }
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const uint32_t k16 = hh ? (mm >> 16) : (mm & 0xFFFFu);
                                const int c = 4 * lane + 2 * jj + hh;
                                if (c < rows)
                                    atomicMin(&sm.colmin[tile][s & 1][c], ((k16 >> 5) << KEY_IDX_BITS) + rbase + (k16 & 31u));
                            }
                        }
                        __syncwarp();  // the scratch is rewritten by the next stage
                        // the 4 warps of this query tile merge: one global atomicMin per train row, tile and stage
                        if (tile == 0)
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                        else
                            asm volatile("bar.sync 2, 128;" ::: "memory");
                        for (int c = row_in_tile; c < rows; c += 128) {
                            atomicMin(p.col_keys + t_row0 + tb + s * NT + c, sm.colmin[tile][s & 1][c]);
                            sm.colmin[tile][s & 1][c] = KEY_NONE;  // reused in stage s + 2, after the barrier of s + 1
                        }
                    }
                    continue;
                }
                // read every 64-column chunk that holds a valid train row, then hand the accumulator back to
                // the tensor core BEFORE folding: the MMAs of the next stage never wait for this warp's arithmetic
                uint32_t vv[G::NCH][32];
#pragma unroll
                for (int h = 0; h < G::NCH; ++h)
                    if ((G::ESPLIT == 1 || h == chunk) && (h == 0 || 64 * h < rows))
                        tmem_ld64_packed(tmem + lane_base + acc * NT + 64 * h, vv[h]);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&sm.d_empty[acc]);
#pragma unroll
                for (int h = 0; h < G::NCH; ++h) {
                    if (G::ESPLIT == 2 && h != chunk) continue;   // warp-uniform
                    if (h > 0 && 64 * h >= rows) break;  // warp-uniform
                    uint32_t (&v)[32] = vv[h];
                    if (COL && !G::REDUX && !G::SCR2) {
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
                            const uint32_t k0 = G::IDXK ? v[j] : add_imad(v[j], one, ((2u * j + 1u) << 16) | (2u * j));
                            const uint32_t k1 = G::IDXK ? v[j + 1] : add_imad(v[j + 1], one, ((2u * j + 3u) << 16) | (2u * j + 2u));
                            m = __vimin3_u16x2(m, k0, k1);
                        }
                        const uint32_t k16 = min(m & 0xFFFFu, m >> 16);
                        b1 = min(b1, ((k16 >> 7) << KEY_IDX_BITS) + (G::IDXK ? jstage : jh) + (k16 & 127u));
                    } else {
                        uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;  // per 16-bit half: best and second best
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t k = G::IDXK ? v[j] : add_imad(v[j], one, ((2u * j + 1u) << 16) | (2u * j));
                            m2 = __vminu2(m2, __vmaxu2(m1, k));
                            m1 = __vminu2(m1, k);
                        }
                        const uint32_t c[4] = {m1 & 0xFFFFu, m1 >> 16, m2 & 0xFFFFu, m2 >> 16};
#pragma unroll
                        for (int i = 0; i < 4; ++i)  // empty slots and invalid train rows (d = 508) are no candidates
                            top2_insert(c[i] >= (505u << 7)
                                            ? KEY_NONE
                                            : ((c[i] >> 7) << KEY_IDX_BITS) + (G::IDXK ? jstage : jh) + (c[i] & 127u),
                                        b1, b2);
                    }
                    if (COL && !G::SCR2) {
                        // ---- column minima over this warp's 32 query rows; lane l ends up with the packed
                        //      minima of train columns 2l, 2l+1 of the chunk: key16 = (d << 5) | row-in-warp ----
                        uint32_t m = 0xFFFFFFFFu;
                        if (G::REDUX) {
                            // acc = d << 7, so acc >> 2 per half (bits 0-1 of every half are zero: both halves
                            // shift together).  redux.min over the packed word is exact for its HIGH half.
                            const uint32_t lane2 = (static_cast<uint32_t>(lane) << 16) | lane;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const uint32_t k = add_imad(v[j] >> 2, one, lane2);
                                const uint32_t hi = __reduce_min_sync(0xFFFFFFFFu, k);
                                const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, __byte_perm(k, 0, 0x1032));
                                if (lane == j) m = __byte_perm(hi, lo, 0x3276);  // [hi.hi16 | lo.hi16]
                            }
                        } else {
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < 32; r += 2) {
                                uint32_t x0, x1;
                                const uint32_t a0 = scr_addr + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2));
                                const uint32_t a1 =
                                    scr_addr + (r + 1) * 128 + ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2));
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(a0) : "memory");
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x1) : "r"(a1) : "memory");
                                const uint32_t k0 = col_key(x0, (static_cast<uint32_t>(r) << 16) | r);
                                const uint32_t k1 = col_key(x1, (static_cast<uint32_t>(r + 1) << 16) | (r + 1));
                                m = __vimin3_u16x2(m, k0, k1);
                            }
                            __syncwarp();  // the scratch is rewritten by the next chunk / stage
                        }
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const uint32_t k16 = hh ? (m >> 16) : (m & 0xFFFFu);
                            const int c = 64 * h + 2 * lane + hh;
                            if (c < rows)
                                atomicMin(&sm.colmin[tile][s & 1][c], ((k16 >> 5) << KEY_IDX_BITS) + rbase + (k16 & 31u));
                        }
                    }
                }
                if (COL && G::SCR2) {
                    // both chunks of the stage through their own scratch: 16 STS.128, one __syncwarp, 64 LDS.32 feeding
                    // two independent minimum chains, one __syncwarp
                    const int nh = rows > 64 ? 2 : 1;   // warp-uniform
#pragma unroll
                    for (int h = 0; h < G::NCH; ++h) {
                        if (h < nh) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t a = scr_addr + h * 4096 + lane * 128 + ((i ^ (lane & 7)) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(vv[h][4 * i]),
                                             "r"(vv[h][4 * i + 1]), "r"(vv[h][4 * i + 2]), "r"(vv[h][4 * i + 3])
                                             : "memory");
                            }
                        }
                    }
                    __syncwarp();
                    uint32_t mc[G::NCH];
#pragma unroll
                    for (int h = 0; h < G::NCH; ++h) mc[h] = 0xFFFFFFFFu;
#pragma unroll
                    for (int r = 0; r < 32; r += 2) {
#pragma unroll
                        for (int h = 0; h < G::NCH; ++h) {
                            if (h < nh) {
                                uint32_t x0, x1;
                                const uint32_t a0 =
                                    scr_addr + h * 4096 + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2));
                                const uint32_t a1 = scr_addr + h * 4096 + (r + 1) * 128 +
                                                    ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2));
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(a0) : "memory");
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x1) : "r"(a1) : "memory");
                                const uint32_t k0 = add_imad(x0 >> 2, one, (static_cast<uint32_t>(r) << 16) | r);
                                const uint32_t k1 = add_imad(x1 >> 2, one, (static_cast<uint32_t>(r + 1) << 16) | (r + 1));
                                mc[h] = __vimin3_u16x2(mc[h], k0, k1);
                            }
                        }
                    }
                    __syncwarp();  // the scratch is rewritten by the next stage
#pragma unroll
                    for (int h = 0; h < G::NCH; ++h) {
                        if (h < nh) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const uint32_t k16 = hh ? (mc[h] >> 16) : (mc[h] & 0xFFFFu);
                                const int c = 64 * h + 2 * lane + hh;
                                if (c < rows)
                                    atomicMin(&sm.colmin[tile][s & 1][c], ((k16 >> 5) << KEY_IDX_BITS) + rbase + (k16 & 31u));
                            }
                        }
                    }
                }
                if (COL && G::ESPLIT == 2) {
                    // the 4 warps of this (query tile, chunk) merge their 64 columns
                    if (64 * chunk < rows) {   // uniform over the group
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * tile + chunk) : "memory");
                        const int c = 64 * chunk + row_in_tile;
                        if (row_in_tile < 64 && c < rows) {
                            atomicMin(p.col_keys + t_row0 + tb + s * NT + c, sm.colmin[tile][s & 1][c]);
                            sm.colmin[tile][s & 1][c] = KEY_NONE;
                        }
                    }
                } else if (COL) {
                    // the 4 warps of this query tile merge: one global atomicMin per train row, tile and stage
                    if (tile == 0)
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    else
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                    for (int c = row_in_tile; c < rows; c += 128) {
                        atomicMin(p.col_keys + t_row0 + tb + s * NT + c, sm.colmin[tile][s & 1][c]);
                        sm.colmin[tile][s & 1][c] = KEY_NONE;  // reused in stage s + 2, after the barrier of s + 1
                    }
                }
            }
            if constexpr (G::LD16) {
                // the four threads (q, 0..3) hold partial results for rows q + 8 ri: merge over m, lane (q, 0) writes
#pragma unroll
                for (int ri = 0; ri < 4; ++ri) {
#pragma unroll
                    for (int d = 1; d <= 2; d <<= 1) {
                        const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, rb1[ri], d);
                        const uint32_t o2 = __shfl_xor_sync(0xFFFFFFFFu, rb2[ri], d);
                        const uint32_t n2 = min(max(rb1[ri], o1), min(rb2[ri], o2));
                        rb1[ri] = min(rb1[ri], o1);
                        rb2[ri] = n2;
                    }
                    const int row4 = qt0 + tile * MQ + quarter * 32 + (lane >> 2) + 8 * ri;
                    if ((lane & 3) == 0 && row4 < nq) {
                        const size_t orow = static_cast<size_t>(p.row_out_off ? p.row_out_off[prob] : q_row0) + row4;
                        if (!TOP2 && p.compact) {
                            uint32_t *c = reinterpret_cast<uint32_t *>(p.row_keys) + orow;
                            if (gridDim.y == 1)
                                *c = rb1[ri];
                            else
                                atomicMin(c, rb1[ri]);
                        } else {
                            uint2 *g = p.row_keys + orow;
                            if (gridDim.y == 1)
                                *g = make_uint2(rb1[ri], TOP2 ? rb2[ri] : KEY_NONE);
                            else if (TOP2)
                                merge_row_keys(g, rb1[ri], rb2[ri]);
                            else
                                atomicMin(&g->x, rb1[ri]);  // second key stays KEY_NONE (pre-set)
                        }
                    }
                }
            } else if (row < nq) {
                const size_t orow = static_cast<size_t>(p.row_out_off ? p.row_out_off[prob] : q_row0) + row;
                const bool sole = gridDim.y == 1 && G::ESPLIT == 1;   // nobody else holds results for this row
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
    } else if (warp < MMA_WARP) {
        // ============================== expander warps ==============================
        constexpr int WPT = W / G::EXP_SPLIT;                // input words per thread
        const int e = tid - N_EPI_WARPS * 32;
        const int r = e % NT;                                // this thread's train row within every stage
        const int k_lo = (e / NT) * WPT;                     // ... and its share of the row's 16 input words
        uint32_t mul[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) mul[s] = (1u << (7 - s)) * static_cast<uint32_t>(p.desc_bytes > 0);
        const uint32_t last_mask = word_mask(p.desc_bytes, ws);   // the word that holds the descriptor's tail + the spare byte
        // Two copies of the stage loop, chosen once per kernel: the 64-byte-row path costs the general one
        // registers and scheduling freedom when both live in the same loop (-3.7 % on the dense sweep).
        auto sweep = [&](auto rows64_c) {
            constexpr bool ROWS64 = decltype(rows64_c)::value;
            for (int s = 0; s < n_stage; ++s) {
                const int b = s % NB;
                const uint32_t ph = (s / NB) & 1;
                const int rows = stage_rows(s), trows = stage_tma_rows(s);
                if (trows > 0) mbar_wait(&sm.raw_full[b], ph);
                mbar_wait(&sm.b_empty[b], ph ^ 1);
                uint32_t w[WPT];
                // 64-byte rows (the compacted features of the sequence pipeline): word k of 32 consecutive rows lives
                // in TWO banks, so the word-by-word read below is a 16-way conflict (this launch ran at 0.45 of the
                // MMA issue floor where 61-byte rows of the same size ran at 0.65).  Read 16-byte chunks instead,
                // each thread starting at chunk (r / 2) mod 4: the four rows of equal parity in a quarter-warp then
                // hit four different bank groups, and the chunk's words go straight to their K positions.
                if (ROWS64 && r < trows && r < rows) {
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
                            const uint32_t s7 = (G::IDXK && tail && j == 3) ? (static_cast<uint32_t>(r) << 24) : 0u;
                            *reinterpret_cast<uint4 *>(d + LBO) = make_uint4(spread80(ww[j], mul[4]), spread80(ww[j], mul[5]),
                                                                             spread80(ww[j], mul[6]), spread80(ww[j], mul[7]) | s7);
                        }
                    }
                    fence_async_smem();
                    mbar_arrive(&sm.b_full[b]);
                    continue;
                }
                if (r >= rows) {
    #pragma unroll
                    for (int k = 0; k < WPT; ++k) w[k] = 0;
                } else if (r < trows) {
                    const int o = r * p.t_stride + 4 * k_lo;  // any alignment: LDS.32 + funnel shift
                    const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(sm.raw[b]) + (o >> 2);
                    const int sh = (o & 3) * 8;
                    uint32_t lo = raw32[0];
    #pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        // words past the descriptor hold stale bytes of the staging buffer (never past its end:
                        // RAW_STAGE has 16 bytes of slack); they are masked / skipped below
                        const uint32_t hi = raw32[k + 1];
                        w[k] = __funnelshift_r(lo, hi, sh);
                        lo = hi;
                    }
                } else {  // rows the bulk copy could not take (misaligned source / tail): read global memory
                    uint32_t wf[W];
                    load_desc_global(stage_src(s) + static_cast<size_t>(r) * p.t_stride, p.desc_bytes, wf);
    #pragma unroll
                    for (int k = 0; k < WPT; ++k) w[k] = wf[(G::EXP_SPLIT == 1 ? 0 : k_lo) + k];
                }
                uint8_t *dst = sm.b[b] + r * 16;
                // valid rows carry 0x80 on the four popc positions, invalid rows on the four marker positions
                const uint32_t spare_lo = (r < rows) ? 0x80000000u : 0u;
                // IDXK: invalid rows carry 0xFF on marker positions 4 and 5, every row its index on position 7
                const uint32_t spare_hi = (r < rows) ? 0u : (G::IDXK ? 0xFF000000u : 0x80000000u);
                auto put = [&](int kk, uint32_t wk, uint32_t slo, uint32_t shi) {
                    uint8_t *d = dst + 2 * kk * LBO;
                    const uint32_t s6 = G::IDXK ? 0u : shi;
                    const uint32_t s7 = G::IDXK ? (kk == ws ? (static_cast<uint32_t>(r) << 24) : 0u) : shi;
                    *reinterpret_cast<uint4 *>(d) = make_uint4(spread80(wk, mul[0]) | slo, spread80(wk, mul[1]) | slo,
                                                               spread80(wk, mul[2]) | slo, spread80(wk, mul[3]) | slo);
                    *reinterpret_cast<uint4 *>(d + LBO) = make_uint4(spread80(wk, mul[4]) | shi, spread80(wk, mul[5]) | shi,
                                                                     spread80(wk, mul[6]) | s6, spread80(wk, mul[7]) | s7);
                };
                if (n_k == W) {   // descriptors of 60..63 bytes (AKAZE: 61): word 15 is the tail word, no per-word tests
    #pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        const bool tail = (G::EXP_SPLIT == 1) ? (k == W - 1) : (k == WPT - 1 && k_lo + WPT == W);
                        if (tail)
                            put(k_lo + k, w[k] & last_mask, spare_lo, spare_hi);
                        else
                            put(k_lo + k, w[k], 0u, 0u);
                    }
                } else {
    #pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        const int kk = k_lo + k;
                        if (kk < ws)
                            put(kk, w[k], 0u, 0u);
                        else if (kk == ws)
                            put(kk, w[k] & last_mask, spare_lo, spare_hi);
                    }
                }
                fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(&sm.b_full[b]);
            }
        };
        if (G::EXP_SPLIT == 1 && n_k == W && p.t_stride == 64)
            sweep(std::true_type{});
        else
            sweep(std::false_type{});
    } else if (warp == MMA_WARP) {
        // ============================== MMA issue (whole warp, one elected lane issues) ===========
        // This thread does nothing but wait and issue: tcgen05.mma blocks when the tensor pipe's queue is
        // full, so every cycle it spends elsewhere (TMA bookkeeping used to cost ~800 cycles per stage here)
        // is a cycle the queue can run dry.
        if (elect_one()) {
            mbar_wait(&sm.a_ready, 0);
            tc_fence_after();
            // K-step k of a B stage: descriptor start address advances by 2 chunks = 2 * LBO bytes
            const uint64_t desc0 = umma_desc(smem_u32(sm.b[0]), LBO, 128);
            // The (stage, tile) jobs are issued back to back.  tcgen05.mma blocks while the tensor pipe's short
            // queue is full, so the barriers of job j + 1 are waited for in the MIDDLE of issuing job j (12 of 16
            // K-steps queued: the wait and the commits hide behind them) — waiting between jobs left the pipe
            // idle for ~250 cycles per tile (measured with clock64 in this thread: issue 1786 + waits 311 + 470
            // other per 2048 cycles of MMA work).
            const int n_jobs = n_stage * n_tiles;
            auto wait_job = [&](int j) {   // everything job j needs: its B stage (first tile only) and its accumulator
                const int s = j / n_tiles, t = j - s * n_tiles;
                if (t == 0) mbar_wait(&sm.b_full[s % NB], (s / NB) & 1);
                mbar_wait(&sm.d_empty[acc_of(s, t, n_tiles)], (acc_use(s, n_tiles) & 1) ^ 1);
                tc_fence_after();
            };
            wait_job(0);
            for (int j = 0; j < n_jobs; ++j) {
                const int s = j / n_tiles, t = j - s * n_tiles;
                const int b = s % NB, acc = acc_of(s, t, n_tiles);
                const uint64_t desc_b = desc0 + static_cast<uint64_t>(b * (G::B_STAGE >> 4));
                const uint32_t d_addr = tmem + acc * NT, a_addr = tmem + G::TMEM_A + t * 128;
                if (n_k == 16) {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    if (j + 1 < n_jobs) wait_job(j + 1);
#pragma unroll
                    for (int k = 12; k < 16; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, 1);
                } else {
                    for (int k = 0; k < n_k; ++k)
                        umma_i8_ts(d_addr, a_addr + 8 * k, desc_b + static_cast<uint64_t>(k * (2 * LBO >> 4)), G::IDESC, k > 0);
                    if (j + 1 < n_jobs) wait_job(j + 1);
                }
                if (t == n_tiles - 1) umma_commit(&sm.b_empty[b]);
                umma_commit(&sm.d_full[acc]);
            }
        }
    } else {
        // ============================== TMA producer (one elected lane) ===========
        // raw[b] is free again once every expander has arrived on b_full[b] for the stage that used it
        if (elect_one()) {
            for (int s = 0; s < n_stage; ++s) {
                if (s >= NB) mbar_wait_relaxed(&sm.b_full[s % NB], ((s - NB) / NB) & 1, 256);
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
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <class G, bool COL, bool TOP2>
int launch_mma(const HammingParams &p, dim3 grid, cudaStream_t stream)
{
    static bool configured = false;  // per instantiation; racing threads set the same value
    constexpr int smem = static_cast<int>(sizeof(typename G::Smem));
    static_assert(smem <= 227 * 1024, "shared memory per CTA");
    if (!configured) {
        SLAMFE_CUDA_OK(cudaFuncSetAttribute(hamming_mma_kernel<G, COL, TOP2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    hamming_mma_kernel<G, COL, TOP2><<<grid, G::THREADS, smem, stream>>>(p);
    return launch_status();
}

#ifdef SLAMFE_MMA_DEV
#include "hamming_mma_persistent.cuh"   // measured and not shipped; see the headers
#include "hamming_mma_pair.cuh"
#endif

template <class G>
int run_geometry(HammingParams p, int n_problems, int max_nq, int max_nt, bool top2, cudaStream_t stream,
                 bool persistent = false)
{
    // One CTA per SM (all 512 TMEM columns).  When the query tiles alone do not fill the machine the train
    // set is cut into slices (grid.y); their results merge exactly through atomicMin / the CAS pair merge,
    // as in the INT kernel.
    const int sms = sm_count();
    const int stages_total = (max_nt + G::NT - 1) / G::NT;
    const long long ctas = static_cast<long long>((max_nq + G::CQ - 1) / G::CQ) * n_problems;
    int slices = 1;
    if (ctas < 2LL * sms) {
        const int want = static_cast<int>((2LL * sms + ctas - 1) / ctas);
        slices = max(1, min(want, stages_total / 4));
    }
    const int stages_per_slice = (stages_total + slices - 1) / slices;
    p.t_slice = stages_per_slice * G::NT;
    const int n_slices = (stages_total + stages_per_slice - 1) / stages_per_slice;
    const dim3 grid((max_nq + G::CQ - 1) / G::CQ, n_slices, n_problems);
    if (grid.y > 65535u || grid.z > 65535u) return SLAMFE_ERANGE;
#ifdef SLAMFE_MMA_DEV
    if constexpr (G::QT == 2 && !G::REDUX && G::EXP_SPLIT == 1 && !G::SCR2 && !G::LD16 && !G::IDXK && G::ESPLIT == 1) if (persistent) {
        const long long total = static_cast<long long>(grid.x) * grid.y * grid.z;
        if (total > 0x7FFF0000LL) return SLAMFE_ERANGE;
        p.jobs_x = static_cast<int>(grid.x);
        p.jobs_y = static_cast<int>(grid.y);
        p.jobs_total = static_cast<int>(total);
        p.job_counter = next_job_counter();
        if (!p.job_counter) return SLAMFE_EINVAL;
        const int ctas = static_cast<int>(std::min<long long>(sms, total));
        if (p.col_keys)
            return top2 ? launch_mma_persistent<G, true, true>(p, ctas, stream)
                        : launch_mma_persistent<G, true, false>(p, ctas, stream);
        return top2 ? launch_mma_persistent<G, false, true>(p, ctas, stream)
                    : launch_mma_persistent<G, false, false>(p, ctas, stream);
    }
#else
    (void)persistent;
#endif
    if (p.col_keys) return top2 ? launch_mma<G, true, true>(p, grid, stream) : launch_mma<G, true, false>(p, grid, stream);
    return top2 ? launch_mma<G, false, true>(p, grid, stream) : launch_mma<G, false, false>(p, grid, stream);
}

}  // namespace

bool hamming_mma_supports(int desc_bytes) { return desc_bytes >= 1 && desc_bytes <= 63; }

int run_hamming_mma(HammingParams p, int n_problems, int max_nq, int max_nt, bool top2, cudaStream_t stream)
{
#ifdef SLAMFE_MMA_DEV
    // Development build only (-DSLAMFE_MMA_DEV, see scripts/README.md): the other instantiations of the template,
    // chosen per process by SLAMFE_MMA_GEOMETRY for A/B runs (profiles/r02_mma_geometries.log).
    //   0 = shipped, 1 = <2,128> redux column minima, 2 = <1,192> redux, 3 = <2,128> two expander threads per row,
    //   4 = <2,128> with one transpose scratch per 64-column chunk, 5 = 16x256b accumulator loads (rows pre-reduced in
    //   registers), 6 = train index in a spare K position (the accumulator is the row key), 7 = 16 epilogue warps (one 64-column
    //   chunk each)
    static const int geometry = [] {
        const char *v = getenv("SLAMFE_MMA_GEOMETRY");
        return v && *v ? atoi(v) : 0;
    }();
    switch (geometry) {
        case 1: return run_geometry<Geo<2, 128, true>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 2: return run_geometry<Geo<1, 192, true>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 3: return run_geometry<Geo<2, 128, false, 2>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 4: return run_geometry<Geo<2, 128, false, 1, true>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 5: return run_geometry<Geo<2, 128, false, 1, false, true>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 6: return run_geometry<Geo<2, 128, false, 1, false, false, true>>(p, n_problems, max_nq, max_nt, top2, stream);
        case 7: return run_geometry<Geo<2, 128, false, 1, false, false, false, 2>>(p, n_problems, max_nq, max_nt, top2, stream);
        default: break;
    }
#endif
#ifdef SLAMFE_MMA_DEV
    static const bool pair = [] {   // A/B of the two-SM kernel in a development build
        const char *v = getenv("SLAMFE_MMA_PAIR");
        return v && *v && atoi(v) != 0;
    }();
    if (pair) return run_pair(p, n_problems, max_nq, max_nt, top2, stream);
    static const bool persistent = [] {   // the persistent form of the kernel (hamming_mma_persistent.cuh), A/B runs only
        const char *v = getenv("SLAMFE_MMA_PERSISTENT");
        return v && *v && atoi(v) != 0;
    }();
    return run_geometry<Geo<2, 128, false, 1>>(p, n_problems, max_nq, max_nt, top2, stream, persistent);
#else
    return run_geometry<Geo<2, 128, false, 1>>(p, n_problems, max_nq, max_nt, top2, stream);
#endif
}

}  // namespace slamfe

#if defined(SLAMFE_MMA_DEV) && defined(SLAMFE_MMA_PROF)
extern "C" int slamfe_dev_mma_prof(unsigned long long *out, int clear)
{
    if (out && cudaMemcpyFromSymbol(out, slamfe::g_mma_prof, sizeof(slamfe::g_mma_prof)) != cudaSuccess) return -1;
    if (clear) {
        static unsigned long long zero[256][16];
        if (cudaMemcpyToSymbol(slamfe::g_mma_prof, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
