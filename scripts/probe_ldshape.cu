// Development probe: register <-> (TMEM lane, column) mapping of tcgen05.ld.16x256b (plain and .pack::16b) on a B200.
// TMEM is filled through tcgen05.st.32x32b with value = lane * 128 + column (< 2^14), then read back with the shapes
// under test; the host prints, per thread, which (lane, column) every register holds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe_ldshape scripts/probe_ldshape.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t *out, int lane_off)
{
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = tid * 128 + c0 + j;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tb + ((uint32_t)(warp * 32) << 16) + c0),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tb + ((uint32_t)(warp * 32 + lane_off) << 16);
    // variant 0: 16x256b.x2 plain: 8 registers
    {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) out[(0 * 128 + tid) * 8 + j] = r[j];
    }
    // variant 1: 16x256b.x2.pack::16b: 8 registers, twice the columns?
    {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) out[(1 * 128 + tid) * 8 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main()
{
    uint32_t *d;
    CK(cudaMalloc(&d, 2 * 128 * 8 * 4));
    for (int lane_off = 0; lane_off <= 16; lane_off += 16) {
        CK(cudaMemset(d, 0xEE, 2 * 128 * 8 * 4));
        probe<<<1, 128>>>(d, lane_off);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h(2 * 128 * 8);
        CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        for (int variant = 0; variant < 2; ++variant) {
            printf("== lane_off %d, %s\n", lane_off, variant ? "16x256b.x2.pack::16b (lo half | hi half)" : "16x256b.x2");
            for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 33, 64}) {
                printf("  thread %3d:", t);
                for (int j = 0; j < 8; ++j) {
                    const uint32_t v = h[(variant * 128 + t) * 8 + j];
                    if (!variant) printf("  r%d=(l%u,c%u)", j, v / 128, v % 128);
                    else printf("  r%d=(l%u,c%u|l%u,c%u)", j, (v & 0xFFFF) / 128, (v & 0xFFFF) % 128, (v >> 16) / 128, (v >> 16) % 128);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
