// Parameter block and small device helpers shared by the two Hamming matcher kernels
// (hamming.cu: INT-pipe carry-save kernel; hamming_mma.cu: tcgen05 tensor-core kernel).
#pragma once
#include "common.cuh"
#include "hamming_core.cuh"

namespace slamfe {

struct HammingParams {
    const uint8_t *q;
    const uint8_t *t;
    int q_stride, t_stride, desc_bytes;
    const int32_t *q_off, *q_cnt, *t_off, *t_cnt;  // null q_off / t_off => single problem
    const int32_t *row_out_off;                    // null => row results are indexed like q rows
    int nq, nt;                                    // single-problem sizes
    int t_index_base;
    int t_slice;      // train rows per blockIdx.y slice (multiple of TS)
    int tma_quantum;  // rows per 16-byte-multiple chunk of the train layout
    uint2 *row_keys;
    uint32_t *col_keys;
    int compact;      // best-only results as one u32 per query row instead of (best, second)
    // persistent tcgen05 kernel only: the (query tile, train slice, problem) job space and the launch's job counter
    int jobs_x, jobs_y, jobs_total;
    uint32_t *job_counter;
};

// Load one descriptor (desc_bytes useful bytes) from global memory into 16 zero-padded words.
__device__ __forceinline__ void load_desc_global(const uint8_t *src, int desc_bytes, uint32_t (&w)[W])
{
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int rem = desc_bytes - 4 * k;
            uint32_t v = 0;
            if (rem >= 4) {
                v = __ldg(s32 + k);
            } else if (rem > 0) {
                for (int b = 0; b < rem; ++b) v |= static_cast<uint32_t>(__ldg(src + 4 * k + b)) << (8 * b);
            }
            w[k] = v;
        }
    } else {
        // misaligned row (cv2's 61-byte stride): aligned 32-bit loads of the words that hold at least one
        // descriptor byte + funnel shifts (64 byte loads per row made the CTA prologue latency-bound).  A word
        // is only read when it contains a valid byte, so nothing outside the allocation's granule is touched.
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
        const int sh = static_cast<int>(a & 3) * 8;
        const int n_need = (static_cast<int>(a & 3) + desc_bytes + 3) >> 2;   // aligned words holding descriptor bytes
        uint32_t lo = __ldg(s32);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const uint32_t hi = (k + 1 < n_need) ? __ldg(s32 + k + 1) : 0u;
            const int rem = desc_bytes - 4 * k;
            const uint32_t m = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u));
            w[k] = __funnelshift_r(lo, hi, sh) & m;
            lo = hi;
        }
    }
}

__device__ __forceinline__ uint32_t word_mask(int desc_bytes, int k)
{
    const int rem = desc_bytes - 4 * k;
    return rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u));
}

// Merge a CTA-local sorted pair (k1 <= k2) into the global (key1, key2) of one query row.
__device__ __forceinline__ void merge_row_keys(uint2 *g, uint32_t k1, uint32_t k2)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(g);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        const uint32_t o1 = static_cast<uint32_t>(assumed), o2 = static_cast<uint32_t>(assumed >> 32);
        const uint32_t n1 = min(o1, k1);
        const uint32_t n2 = min(max(o1, k1), min(o2, k2));
        const unsigned long long nv = (static_cast<unsigned long long>(n2) << 32) | n1;
        if (nv == assumed) break;
        old = atomicCAS(a, assumed, nv);
    } while (old != assumed);
}

// Entry of the tensor-core kernel (hamming_mma.cu); same contract as run_hamming's launch: the key
// tables were pre-set to KEY_NONE by the caller.
int run_hamming_mma(HammingParams p, int n_problems, int max_nq, int max_nt, bool top2, cudaStream_t stream);
// desc_bytes the tensor-core kernel handles (it needs a spare descriptor byte inside the 64-byte K range)
bool hamming_mma_supports(int desc_bytes);

}  // namespace slamfe
