// Development probe (not part of libslamfe): tcgen05.mma.cta_group::2 kind::i8 on a CTA pair, the facts a
// two-SM form of the matcher would rest on, pinned on a real B200 before any kernel is written.
//   check : D (256 x N) = A (256 x K) . B (N x K)^T with A in TMEM (128 rows per CTA), B in shared memory (K-major
//           SWIZZLE_NONE core matrices, N/2 rows per CTA), issued by one thread of the leader CTA, completion
//           multicast to one mbarrier in each CTA; readiness of the peer signalled by a REMOTE mbarrier arrive.
//           Reports which reading of "half of B per CTA" matches the host reference.
//   rate  : cycles per M256 N128 K32 instruction at the issue floor (one CTA pair per SM pair)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_2cta probe_2cta.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// wait on a barrier of this CTA that is also arrived on from the peer CTA / by a multicast commit
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank)
{
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// completion of all MMAs issued so far -> the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t *bar)
{
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
#define LD32(taddr, v)                                                                                               \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31}, [%32];"                                                                                       \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
        : "r"(taddr)                                                                                                 \
        : "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// CTA r of the pair: A rows [128 r, 128 r + 128) in its TMEM, B rows [r N/2, (r + 1) N/2) in its shared memory.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
    mma_check2(const uint8_t *A, const uint8_t *B, uint32_t *D, int K, int N, uint32_t idesc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done_bar, peer_bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t rank = cta_rank();
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NH = N / 2;
    uint8_t *sB = smem;
    for (int i = tid; i < NH * K / 16; i += 128) {
        const int r = i % NH, c = i / NH;
        *reinterpret_cast<uint4 *>(sB + c * (NH * 16) + r * 16) =
            *reinterpret_cast<const uint4 *>(B + (size_t)(rank * NH + r) * K + c * 16);
    }
    fence_async_smem();
    if (tid == 0) {
        mbar_init(&done_bar, 1);
        mbar_init(&peer_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc2(&tmem_base_s, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t a_tmem = tb + 256;
    for (int k = 0; k < K / 32; ++k) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j)
            v[j] = *reinterpret_cast<const uint32_t *>(A + (size_t)(rank * 128 + tid) * K + k * 32 + j * 4);
        st8(a_tmem + ((uint32_t)(warp * 32) << 16) + k * 8, v);
    }
    wait_st();
    fence_before();
    __syncthreads();   // this CTA's operands are in place
    cluster_sync();    // barriers of both CTAs are initialised before anybody arrives remotely
    if (rank == 1 && tid == 0) mbar_arrive_remote(&peer_bar, 0);   // "the peer's operands are in place"
    if (rank == 0 && tid == 0) {
        mbar_wait_cluster(&peer_bar, 0);
        fence_after();
        const uint32_t lboB = NH * 16, sbo = 128;
        for (int k = 0; k < K / 32; ++k) {
            const uint64_t db = make_desc(smem_u32(sB) + k * 2 * lboB, lboB, sbo);
            mma2_ts(tb, a_tmem + k * 8, db, idesc, k > 0);
        }
        umma_commit2(&done_bar);
    }
    mbar_wait_cluster(&done_bar, 0);
    fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        LD32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        wait_ld();
        for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + tid) * N + c0 + j] = v[j];
    }
    fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tb, 512);
}

// rate: `groups` x 16 MMAs (M256 N128 K32, A in TMEM) issued back to back by the leader, commit per group
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_rate2(int N, uint32_t idesc, int groups, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t rank = cta_rank();
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * (N / 2) * 16 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x80808080u, 0, 0x80u, 0);
    fence_async_smem();
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc2(&tmem_base_s, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    uint32_t v[8] = {0x01010101u, 0xFFFFFFFFu, 0x01FF01FFu, 0, 1, 2, 3, 4};
    for (int k = 0; k < 16; ++k) st8(tb + 256 + ((uint32_t)(warp * 32) << 16) + k * 8, v);
    wait_st();
    fence_before();
    __syncthreads();
    cluster_sync();
    if (rank == 0 && tid == 0) {
        fence_after();
        const uint32_t lbo = (N / 2) * 16;
        const uint64_t d0 = make_desc(smem_u32(smem), lbo, 128);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tb + (N == 128 ? (g & 1) * 128 : 0);
#pragma unroll
            for (int k = 0; k < 16; ++k) mma2_ts(d, tb + 256 + 8 * k, d0 + (uint64_t)(k * (2 * lbo >> 4)), idesc, k > 0);
        }
        umma_commit2(&bar);   // one arrival in each CTA once every MMA above is complete
        mbar_wait_cluster(&bar, 0);
        cycles[blockIdx.x / 2] = clock64() - t0;
    } else if (tid == 0) {
        mbar_wait_cluster(&bar, 0);
    }
    fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tb, 512);
}

static uint32_t make_idesc(int M, int N)
{   // D = S32, A signed 8-bit, B unsigned 8-bit, both K-major
    return (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static uint32_t rng_state = 12345u;
static uint32_t rnd()
{
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}

static int run_check(int K, int N)
{
    std::vector<uint8_t> A(256 * K), B((size_t)N * K);
    for (auto &a : A) a = (uint8_t)(int8_t)((int)(rnd() % 255) - 127);
    for (auto &b : B) b = (uint8_t)(rnd() % 256);
    uint8_t *dA, *dB;
    uint32_t *dD;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, (size_t)256 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xEE, (size_t)256 * N * 4));
    const int smem = (N / 2) * K;
    CK(cudaFuncSetAttribute(mma_check2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_check2<<<2, 128, smem>>>(dA, dB, dD, K, N, make_idesc(256, N));
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> D((size_t)256 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // reading 0: CTA r's B rows are D columns [r N/2, (r+1) N/2) (what the kernel loaded);
    // reading 1: the halves are swapped; reading 2: each CTA only sees its own half (columns repeat)
    long bad[3] = {0, 0, 0};
    int shown = 0;
    for (int r = 0; r < 256; ++r)
        for (int c = 0; c < N; ++c) {
            const int32_t got = (int32_t)D[(size_t)r * N + c];
            for (int reading = 0; reading < 3; ++reading) {
                int cb = c;
                if (reading == 1) cb = (c + N / 2) % N;
                if (reading == 2) cb = (r / 128) * (N / 2) + c % (N / 2);
                long ref = 0;
                for (int k = 0; k < K; ++k) ref += (long)(int8_t)A[(size_t)r * K + k] * (long)B[(size_t)cb * K + k];
                if (got != ref) {
                    ++bad[reading];
                    if (reading == 0 && shown < 4) { printf("   mismatch r=%d c=%d got=%d ref=%ld\n", r, c, got, ref); ++shown; }
                }
            }
        }
    printf("check2 K=%d N=%d : reading0 (rank r = columns r N/2..) %s (%ld), reading1 (swapped) %ld, reading2 (own half only) %ld of %d\n",
           K, N, bad[0] ? "FAIL" : "OK", bad[0], bad[1], bad[2], 256 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad[0] != 0;
}

static void run_rate(int N, int pairs)
{
    long long *dc;
    CK(cudaMalloc(&dc, pairs * sizeof(long long)));
    const int smem = 32 * (N / 2) * 16;
    CK(cudaFuncSetAttribute(mma_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int groups = 64;
    mma_rate2<<<2 * pairs, 128, smem>>>(N, make_idesc(256, N), groups, dc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(pairs);
    CK(cudaMemcpy(c.data(), dc, pairs * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0, mn = 1ll << 60;
    for (auto v : c) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    printf("rate2 M256 N%d K32 cta_group::2, %d pairs: %.1f .. %.1f cycles per MMA (%.2f .. %.2f descriptor pairs/clk/SM at K=512)\n", N,
           pairs, (double)mn / (groups * 16), (double)mx / (groups * 16), 128.0 * N / ((double)mx / groups),
           128.0 * N / ((double)mn / groups));
    cudaFree(dc);
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "check";
    if (!strcmp(mode, "check")) {
        int rc = run_check(argc > 2 ? atoi(argv[2]) : 64, argc > 3 ? atoi(argv[3]) : 128);
        return rc;
    }
    if (!strcmp(mode, "rate")) {
        run_rate(128, 1);
        run_rate(128, 74);
        run_rate(256, 74);
        return 0;
    }
    return 1;
}
