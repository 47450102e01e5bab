// Host/device function qualifiers shared by the arithmetic headers that the CPU test suite also
// compiles for the host (p3p.cuh, ransac_core.cuh).
#pragma once
#ifdef __CUDACC__
#define SLAMFE_HD __host__ __device__ __forceinline__
#define SLAMFE_HD_PLAIN __host__ __device__ inline
#define SLAMFE_HD_NOINLINE __host__ __device__ __noinline__ inline
#else
#define SLAMFE_HD inline
#define SLAMFE_HD_PLAIN inline
#define SLAMFE_HD_NOINLINE inline
#endif
