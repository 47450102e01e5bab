// Pipe-peak micro-benchmarks and library utilities — sm_100a.
//
// MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks; the matcher is bound by the
// INT/popc pipe and the RANSAC scorer by the fp64 FMA pipe, so the denominators of their
// rooflines are measured here, on the box, with register-only kernels.
#include "common.cuh"

namespace slamfe {

int sm_count()
{
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

namespace {

constexpr int PEAK_UNROLL = 16;

// mode 0: 16 independent POPC chains per thread (pure popc issue rate).
__global__ void peak_popc_kernel(int iters, uint32_t *sink)
{
    uint32_t x[PEAK_UNROLL], acc[PEAK_UNROLL];
#pragma unroll
    for (int k = 0; k < PEAK_UNROLL; ++k) {
        x[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
        acc[k] = 0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_UNROLL; ++k) {
            acc[k] = __popc(x[k] ^ acc[k]) + it;  // popc result feeds the next popc: no hoisting
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_UNROLL; ++k) s += acc[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 1: the matcher's instruction mix per descriptor pair: 16 x (XOR + POPC) + add tree +
// key build + top-2 update, operands in registers (no shared memory, no global traffic).
__global__ void peak_matcher_mix_kernel(int iters, uint32_t *sink)
{
    uint32_t q[16], t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        q[k] = threadIdx.x * 2654435761u + k * 40503u;
        t[k] = blockIdx.x * 97u + k;
    }
    uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
    for (int it = 0; it < iters; ++it) {
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) d += __popc(q[k] ^ t[k]);
        const uint32_t key = (d << 22) + it;
        const uint32_t hi = max(b1, key);
        b1 = min(b1, key);
        b2 = min(b2, hi);
#pragma unroll
        for (int k = 0; k < 16; ++k) t[k] += 0x9E3779B9u;  // next "train row" (extra IADD, counted as overhead)
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = b1 ^ b2;
}

// mode 2: 8 independent fp64 FMA chains per thread.
__global__ void peak_fp64_kernel(int iters, uint32_t *sink)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<uint32_t>(__double2ll_rn(s));
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_version(void) { return SLAMFE_ABI_VERSION; }

extern "C" const char *slamfe_error_string(int code)
{
    if (code == 0) return "ok";
    if (code == SLAMFE_EINVAL) return "slamfe: invalid argument";
    if (code == SLAMFE_ERANGE) return "slamfe: size out of range for key encoding / grid";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "slamfe: unknown error";
}

extern "C" int slamfe_peak_kernel(int mode, int iters, int grid, int block, uint32_t *sink, int *ops_per_thread_iter,
                                  slamfe_stream_t stream)
{
    if (iters <= 0 || grid <= 0 || block <= 0 || block > 1024 || !sink) return SLAMFE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int ops = 0;
    switch (mode) {
        case 0:
            peak_popc_kernel<<<grid, block, 0, s>>>(iters, sink);
            ops = PEAK_UNROLL;  // popc per thread per iteration
            break;
        case 1:
            peak_matcher_mix_kernel<<<grid, block, 0, s>>>(iters, sink);
            ops = 16;  // popc per thread per iteration (= one descriptor pair)
            break;
        case 2:
            peak_fp64_kernel<<<grid, block, 0, s>>>(iters, sink);
            ops = 8;  // fp64 FMA per thread per iteration
            break;
        default:
            return SLAMFE_EINVAL;
    }
    if (ops_per_thread_iter) *ops_per_thread_iter = ops;
    return launch_status();
}
