// Development probe for the tcgen05 matcher (not part of libslamfe): pins the facts the kernel
// design rests on, on a real B200, before the kernel is written.
//   check  : tcgen05.mma kind::i8 / kind::f8f6f4 with hand-built K-major SWIZZLE_NONE shared-memory
//            descriptors (and A from TMEM) against a host reference -> which LBO/SBO reading is right
//   f8pow2 : mixed e5m2 x e4m3 products 2^e * 2^-e (incl. e4m3 subnormals) accumulate exactly, and
//            the 2^23 offset pad makes the fp32 accumulator bits an integer
//   rate   : cycles per MMA instruction (M=128, N=128/256, SS / TS, i8 / f8)
//   ldtm   : tcgen05.ld 32x32b.x32 throughput with 4 / 8 warps
//   redux  : redux.sync.min.u32 throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tcgen05 probe_tcgen05.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes, contiguous (128 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;  // descriptor version (Blackwell)
    return d;
}
template <int KIND>  // 0 = i8, 1 = f8f6f4
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
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

// ------------------------------------------------------------------------------------------------
// check: D (128 x N) = A (128 x K bytes) * B (N x K bytes)^T
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) mma_check(const uint8_t *A, const uint8_t *B, uint32_t *D, int K, int N,
                                                 uint32_t idesc, int ts, int swap)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sA = smem, *sB = smem + 128 * K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // canonical layout: 16-byte K chunk c, row r -> c * (rows * 16) + r * 16
    for (int i = tid; i < 128 * K / 16; i += 128) {
        const int r = i % 128, c = i / 128;
        *reinterpret_cast<uint4 *>(sA + c * 2048 + r * 16) = *reinterpret_cast<const uint4 *>(A + (size_t)r * K + c * 16);
    }
    for (int i = tid; i < N * K / 16; i += 128) {
        const int r = i % N, c = i / N;
        *reinterpret_cast<uint4 *>(sB + c * (N * 16) + r * 16) = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + c * 16);
    }
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t a_tmem = tb + 256;  // columns 256.. hold A in TS mode (K/4 columns)
    if (ts) {
        // thread t owns TMEM lane t = row t of A; 32 bytes of K = 8 columns per MMA step
        for (int k = 0; k < K / 32; ++k) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t *>(A + (size_t)tid * K + k * 32 + j * 4);
            st8(a_tmem + ((uint32_t)(warp * 32) << 16) + k * 8, v);
        }
        wait_st();
        fence_before();
        __syncthreads();
        fence_after();
    }
    if (tid == 0) {
        const uint32_t lboA = 2048, lboB = N * 16, sbo = 128;
        for (int k = 0; k < K / 32; ++k) {
            const uint64_t da = swap ? make_desc(smem_u32(sA) + k * 2 * lboA, sbo, lboA)
                                     : make_desc(smem_u32(sA) + k * 2 * lboA, lboA, sbo);
            const uint64_t db = swap ? make_desc(smem_u32(sB) + k * 2 * lboB, sbo, lboB)
                                     : make_desc(smem_u32(sB) + k * 2 * lboB, lboB, sbo);
            if (ts)
                mma_ts<KIND>(tb, a_tmem + k * 8, db, idesc, k > 0);
            else
                mma_ss<KIND>(tb, da, db, idesc, k > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        LD32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        wait_ld();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------
// rate: one CTA per SM, thread 0 issues `iters` MMAs back to back
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) mma_rate(int N, uint32_t idesc, int ts, int iters, long long *cycles, int nacc = 1)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 512 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    if (tid == 0) {
        const uint32_t sA = smem_u32(smem), sB = sA + 128 * 512;
        const uint32_t lboA = 2048, lboB = N * 16, sbo = 128;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int k = it & 15;
            const uint64_t da = make_desc(sA + k * 2 * lboA, lboA, sbo);
            const uint64_t db = make_desc(sB + k * 2 * lboB, lboB, sbo);
            const uint32_t dd = tb + ((it / 1) % nacc) * N;   // nacc independent accumulators, round robin
            if (ts)
                mma_ts<KIND>(dd, tb + 256 + k * 8, db, idesc, 1);
            else
                mma_ss<KIND>(dd, da, db, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------
// rate3: the issue loop of the shipped matcher — one elected lane, 16 unrolled TS-mode MMAs per group with
// precomputed descriptors, `nacc` accumulators used round robin per group — to separate the cost of the
// tcgen05.mma itself from the cost of a slow issuing thread (rate / rate2 above)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one_lane()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(128) mma_rate_tight(int N, uint32_t idesc, int groups, int nacc, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * 512 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    if (warp == 1) {
        const bool leader = elect_one_lane();
        const uint32_t lbo = N * 16;
        const uint64_t desc0 = make_desc(smem_u32(smem), lbo, 128);
        const uint32_t a0 = tb + 512 - 128;            // A tile in the last 128 columns
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t dd = tb + (g % nacc) * N;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    mma_ts<0>(dd, a0 + 8 * k, desc0 + (uint64_t)(k * (2 * lbo >> 4)), idesc, k > 0);
            }
            __syncwarp();
        }
        if (leader) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (leader) cycles[blockIdx.x] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

static uint32_t make_idesc(int kind, int a_signed, int b_signed, int M, int N);
static void run_rate_tight(int N, int nacc, int sms)
{
    const int groups = 512;
    long long *dc;
    CK(cudaMalloc(&dc, sizeof(long long) * sms));
    const int smem = N * 512;
    const uint32_t idesc = make_idesc(0, 1, 0, 128, N);
    CK(cudaFuncSetAttribute(mma_rate_tight, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        mma_rate_tight<<<sms, 128, smem>>>(N, idesc, groups, nacc, dc);
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> c(sms);
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : c) mx = v > mx ? v : mx;
    printf("rate3 (tight issue, TS, i8) N=%d nacc=%d : %.1f cycles/MMA, %.2f pairs/clk/SM (K=512)\n", N, nacc,
           (double)mx / (groups * 16), 128.0 * N / ((double)mx / groups));
    cudaFree(dc);
}

// ------------------------------------------------------------------------------------------------
// rate4: the tight MMA loop of rate3 (N = 128, two accumulators) while OTHER warps of the CTA keep a second
// resource busy: mode 1 = tcgen05.ld of accumulator columns (what the epilogue does), mode 2 = STS.128
// streams into shared memory (what the expanders do), mode 3 = both.  What slows the tensor pipe?
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(416) mma_rate_contended(int N, uint32_t idesc, int groups, int mode, int n_ld_warps,
                                                           int n_st_warps, long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < N * 512 / 16; i += 416) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s;
    if (warp == 12) {
        const bool leader = elect_one_lane();
        const uint32_t lbo = N * 16;
        const uint64_t desc0 = make_desc(smem_u32(smem), lbo, 128);
        const uint32_t a0 = tb + 512 - 128;
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t dd = tb + (g & 1) * N;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    mma_ts<0>(dd, a0 + 8 * k, desc0 + (uint64_t)(k * (2 * lbo >> 4)), idesc, k > 0);
            }
            __syncwarp();
        }
        if (leader) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (leader) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
    } else if (warp < n_ld_warps && (mode & 1)) {
        uint32_t acc = 0;
        while (!stop) {
            uint32_t v[32];
            LD32(tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64, v);
            wait_ld();
            for (int j = 0; j < 32; ++j) acc ^= v[j];
        }
        sink[blockIdx.x * 416 + tid] = acc;
    } else if (warp >= 8 && warp < 8 + n_st_warps && (mode & 2)) {
        uint8_t *dst = smem + N * 512 + (warp - 8) * 16384 + lane * 16;   // private 16 KB per warp
        uint32_t x = tid;
        while (!stop) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(dst + k * 512)), "r"(x), "r"(x + 1),
                             "r"(x + 2), "r"(x + 3) : "memory");
            ++x;
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

static void run_rate_contended(int mode, int n_ld, int n_st, int sms)
{
    const int N = 128, groups = 512;
    long long *dc; uint32_t *sink;
    CK(cudaMalloc(&dc, sizeof(long long) * sms));
    CK(cudaMalloc(&sink, 4 * sms * 416));
    const int smem = N * 512 + 4 * 16384;
    const uint32_t idesc = make_idesc(0, 1, 0, 128, N);
    CK(cudaFuncSetAttribute(mma_rate_contended, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        mma_rate_contended<<<sms, 416, smem>>>(N, idesc, groups, mode, n_ld, n_st, dc, sink);
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> c(sms);
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : c) mx = v > mx ? v : mx;
    printf("rate4 N=128 mode=%d (ld warps %d, st warps %d): %.1f cycles/MMA\n", mode, (mode & 1) ? n_ld : 0,
           (mode & 2) ? n_st : 0, (double)mx / (groups * 16));
    cudaFree(dc); cudaFree(sink);
}

// ------------------------------------------------------------------------------------------------
// ldtm: nwarps warps read 32 lanes x 32 columns per instruction, `batch` loads per wait
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ldtm_rate(int iters, int batch, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[4][32];
        if (batch == 4) {
            LD32(tb + 0, v[0]);
            LD32(tb + 32, v[1]);
            LD32(tb + 64, v[2]);
            LD32(tb + 96, v[3]);
            wait_ld();
            acc ^= v[0][0] ^ v[1][1] ^ v[2][2] ^ v[3][3];
        } else {
            LD32(tb + (it & 3) * 32, v[0]);
            wait_ld();
            acc ^= v[0][0];
        }
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

__global__ void __launch_bounds__(512) redux_rate(int iters, long long *cycles, uint32_t *sink)
{
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t x0 = tid * 2654435761u, x1 = x0 ^ 0x1234567u, x2 = x0 + 77u, x3 = x0 * 3u;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        a0 += __reduce_min_sync(0xFFFFFFFFu, x0 + it);
        a1 += __reduce_min_sync(0xFFFFFFFFu, x1 + it);
        a2 += __reduce_min_sync(0xFFFFFFFFu, x2 + it);
        a3 += __reduce_min_sync(0xFFFFFFFFu, x3 + it);
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = a0 ^ a1 ^ a2 ^ a3;
}

// ------------------------------------------------------------------------------------------------
static uint32_t make_idesc(int kind, int afmt, int bfmt, int M, int N)
{
    const uint32_t cfmt = kind == 0 ? 2u : 1u;  // S32 / F32
    return (cfmt << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
static uint32_t rng_state = 12345u;
static uint32_t rnd()
{
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}

static int run_check(int kind, int ts, int swap, int K, int N, int f8pow2)
{
    std::vector<uint8_t> A(128 * K), B((size_t)N * K);
    std::vector<double> Av(128 * K), Bv((size_t)N * K);
    if (kind == 0) {
        for (size_t i = 0; i < A.size(); ++i) {
            int v = (int)(rnd() % 255) - 127;  // s8
            A[i] = (uint8_t)(int8_t)v;
            Av[i] = v;
        }
        for (size_t i = 0; i < B.size(); ++i) {
            int v = rnd() % 256;  // u8
            B[i] = (uint8_t)v;
            Bv[i] = v;
        }
    } else if (!f8pow2) {
        // e4m3 small integers: 0, +-1, +-2 on both sides
        static const uint8_t enc[5] = {0x00, 0x38, 0xB8, 0x40, 0xC0};
        static const double val[5] = {0, 1, -1, 2, -2};
        for (size_t i = 0; i < A.size(); ++i) { int s = rnd() % 5; A[i] = enc[s]; Av[i] = val[s]; }
        for (size_t i = 0; i < B.size(); ++i) { int s = rnd() % 5; B[i] = enc[s]; Bv[i] = val[s]; }
    } else {
        // A e5m2 = +-2^e, B e4m3 = single-bit bytes (1 << p), p = 0..6: 2^-9,-8,-7,-6,-5,-3,+1
        static const int bexp[7] = {-9, -8, -7, -6, -5, -3, 1};
        for (int r = 0; r < 128; ++r)
            for (int k = 0; k < K; ++k) {
                const int p = k % 7;
                const int e = -bexp[p];
                const int sgn = rnd() & 1;
                A[r * K + k] = (uint8_t)((sgn << 7) | ((e + 15) << 2));
                Av[r * K + k] = (sgn ? -1.0 : 1.0) * ldexp(1.0, e);
            }
        for (int r = 0; r < N; ++r)
            for (int k = 0; k < K; ++k) {
                const int p = k % 7;
                const int bit = rnd() & 1;
                B[(size_t)r * K + k] = bit ? (uint8_t)(1u << p) : 0;
                Bv[(size_t)r * K + k] = bit ? ldexp(1.0, bexp[p]) : 0.0;
            }
        // last two K columns: offset pad 2^23 = 32768 (e5m2 0x78) * 256 (e4m3 0x78), and 2 * 256 = 512
        for (int r = 0; r < 128; ++r) {
            A[r * K + K - 1] = 0x78; Av[r * K + K - 1] = 32768.0;
            A[r * K + K - 2] = 0x40; Av[r * K + K - 2] = 2.0;
        }
        for (int r = 0; r < N; ++r) {
            B[(size_t)r * K + K - 1] = 0x78; Bv[(size_t)r * K + K - 1] = 256.0;
            B[(size_t)r * K + K - 2] = 0x78; Bv[(size_t)r * K + K - 2] = 256.0;
        }
    }
    uint8_t *dA, *dB;
    uint32_t *dD;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xEE, (size_t)128 * N * 4));
    const int smem = (128 + N) * K;
    const uint32_t idesc = kind == 0 ? make_idesc(0, 1, 0, 128, N) : make_idesc(1, f8pow2 ? 1 : 0, 0, 128, N);
    if (kind == 0) {
        CK(cudaFuncSetAttribute(mma_check<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mma_check<0><<<1, 128, smem>>>(dA, dB, dD, K, N, idesc, ts, swap);
    } else {
        CK(cudaFuncSetAttribute(mma_check<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mma_check<1><<<1, 128, smem>>>(dA, dB, dD, K, N, idesc, ts, swap);
    }
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> D((size_t)128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    int shown = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += Av[r * K + k] * Bv[(size_t)c * K + k];
            double got;
            const uint32_t bits = D[(size_t)r * N + c];
            if (kind == 0) got = (double)(int32_t)bits;
            else { float f; memcpy(&f, &bits, 4); got = f; }
            if (got != ref) {
                ++bad;
                if (shown < 4) { printf("   mismatch r=%d c=%d got=%g (0x%08x) ref=%g\n", r, c, got, bits, ref); ++shown; }
            }
        }
    if (f8pow2 && !bad) {
        const uint32_t bits = D[5 * N + 7];
        printf("   f8pow2 sample bits 0x%08x (expect 0x4B000000 + 512 + signed sum)\n", bits);
    }
    printf("check kind=%s ts=%d swap=%d K=%d N=%d f8pow2=%d : %s (%ld / %d mismatches)\n", kind ? "f8f6f4" : "i8", ts,
           swap, K, N, f8pow2, bad ? "FAIL" : "OK", bad, 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

static void run_rate(int kind, int ts, int N, int sms, int nacc = 1)
{
    const int iters = 8192;
    long long *dc;
    CK(cudaMalloc(&dc, sizeof(long long) * sms));
    const int smem = (128 + N) * 512;
    const uint32_t idesc = kind == 0 ? make_idesc(0, 1, 0, 128, N) : make_idesc(1, 1, 0, 128, N);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        if (kind == 0) {
            CK(cudaFuncSetAttribute(mma_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            mma_rate<0><<<sms, 128, smem>>>(N, idesc, ts, iters, dc, nacc);
        } else {
            CK(cudaFuncSetAttribute(mma_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            mma_rate<1><<<sms, 128, smem>>>(N, idesc, ts, iters, dc, nacc);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> c(sms);
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : c) mx = v > mx ? v : mx;
    const double macs = (double)sms * iters * 128.0 * N * 32.0;
    printf("rate nacc=%d kind=%s ts=%d N=%d : %.1f cycles/MMA (max CTA), %.3f ms, %.2f P MAC/s, %.2f T desc-pairs/s (K=512)\n",
           nacc, kind ? "f8f6f4" : "i8", ts, N, (double)mx / iters, ms, macs / ms / 1e12, macs / 512.0 / ms / 1e9);
    cudaFree(dc);
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "check";
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, mode %s\n", prop.name, sms, mode);
    if (!strcmp(mode, "check")) {
        const int kind = argc > 2 ? atoi(argv[2]) : 0, ts = argc > 3 ? atoi(argv[3]) : 0, swap = argc > 4 ? atoi(argv[4]) : 0;
        const int K = argc > 5 ? atoi(argv[5]) : 64, N = argc > 6 ? atoi(argv[6]) : 128, p2 = argc > 7 ? atoi(argv[7]) : 0;
        return run_check(kind, ts, swap, K, N, p2);
    }
    if (!strcmp(mode, "rate")) {
        for (int kind = 0; kind < 2; ++kind)
            for (int ts = 0; ts < 2; ++ts)
                for (int N = 128; N <= 256; N += 128) run_rate(kind, ts, N, sms);
        return 0;
    }
    if (!strcmp(mode, "rate2")) {   // i8, TS: N x number of independent accumulators
        for (int N = 32; N <= 256; N *= 2)
            for (int nacc = 1; nacc <= 256 / N && nacc <= 4; nacc *= 2) run_rate(0, 1, N, sms, nacc);
        run_rate(0, 1, 192, sms, 1);
        run_rate(0, 0, 64, sms, 1); run_rate(0, 0, 64, sms, 4);
        return 0;
    }
    if (!strcmp(mode, "rate3")) {
        for (int N = 64; N <= 256; N += 64)
            for (int nacc = 1; nacc <= 2 && nacc * N <= 384; ++nacc) run_rate_tight(N, nacc, sms);
        return 0;
    }
    if (!strcmp(mode, "rate4")) {
        run_rate_contended(0, 0, 0, sms);
        run_rate_contended(1, 4, 0, sms); run_rate_contended(1, 8, 0, sms);
        run_rate_contended(2, 0, 2, sms); run_rate_contended(2, 0, 4, sms);
        run_rate_contended(3, 8, 4, sms);
        return 0;
    }
    if (!strcmp(mode, "ldtm")) {
        long long *dc; uint32_t *sink;
        CK(cudaMalloc(&dc, sizeof(long long) * sms * 8));
        CK(cudaMalloc(&sink, 4 * sms * 256));
        for (int nw = 4; nw <= 8; nw += 4)
            for (int batch = 1; batch <= 4; batch += 3) {
                const int iters = 4096;
                ldtm_rate<<<sms, nw * 32>>>(iters, batch, dc, sink);
                CK(cudaDeviceSynchronize());
                std::vector<long long> c(sms * 8);
                CK(cudaMemcpy(c.data(), dc, sizeof(long long) * sms * 8, cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (int b = 0; b < sms; ++b) for (int w = 0; w < nw; ++w) mx = c[b * 8 + w] > mx ? c[b * 8 + w] : mx;
                const double per_sm_bytes = (double)nw * iters * batch * 32 * 32 * 4;
                printf("ldtm warps=%d batch=%d : %.1f cycles per LD32 per warp, %.1f B/clk/SM\n", nw, batch,
                       (double)mx / (iters * batch), per_sm_bytes / mx);
            }
        return 0;
    }
    if (!strcmp(mode, "redux")) {
        long long *dc; uint32_t *sink;
        CK(cudaMalloc(&dc, sizeof(long long) * sms * 16));
        CK(cudaMalloc(&sink, 4 * sms * 512));
        for (int nw = 4; nw <= 16; nw *= 2) {
            const int iters = 4096;
            redux_rate<<<sms, nw * 32>>>(iters, dc, sink);
            CK(cudaDeviceSynchronize());
            std::vector<long long> c(sms * 16);
            CK(cudaMemcpy(c.data(), dc, sizeof(long long) * sms * 16, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int b = 0; b < sms; ++b) for (int w = 0; w < nw; ++w) mx = c[b * 16 + w] > mx ? c[b * 16 + w] : mx;
            printf("redux warps=%d : %.2f warp-redux per clk per SM (%.2f cycles per redux per warp)\n", nw,
                   (double)nw * iters * 4 / mx, (double)mx / (iters * 4));
        }
        return 0;
    }
    printf("unknown mode\n");
    return 1;
}
