// TEST INFRASTRUCTURE ONLY.  Compiles the product's minimal-solver header (csrc/p3p.cuh, plain C++
// under SLAMFE_HD) for the HOST so that the CPU test suite (-m "not gpu") can pin its arithmetic
// against oracle/p3p_oracle.py, cv2.SOLVEPNP_P3P and ground truth without a GPU.  The product never
// loads this library; on the GPU the same header runs inside csrc/ransac_gen.cu.
#include "../67604-slam---video-navigation_b200/csrc/p3p.cuh"

using namespace slamfe::p3p;

extern "C" int p3p_host_solve(const double *P, const double *uv, const double *K, const double *Kinv, double *T)
{
    Vec3 p[4];
    double px[4][2];
    for (int i = 0; i < 4; ++i) {
        p[i] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        px[i][0] = uv[2 * i];
        px[i][1] = uv[2 * i + 1];
    }
    return solve_sample(p, px, K, Kinv, T) ? 1 : 0;
}

extern "C" int p3p_host_quartic(const double *A, double *roots)
{
    double x[4];
    const int n = quartic_real_roots(A[0], A[1], A[2], A[3], A[4], x);
    for (int k = 0; k < n; ++k) roots[k] = x[k];
    return n;
}

extern "C" void p3p_host_sample4(uint64_t seed, uint32_t frame, uint32_t hyp, int n, int *idx)
{
    int out[4];
    sample4(seed, frame, hyp, n, out);
    for (int k = 0; k < 4; ++k) idx[k] = out[k];
}
