// slamfe_pnp_refit: the final solve of ransac_pnp (final_project/algorithms/ransac.py:185-193,
// cv2.solvePnP(EPNP) on all inliers of the best hypothesis) for every frame pair / loop-closure
// candidate of a batch in ONE launch.  One CTA per problem: the consensus set (best_mask of
// slamfe_ransac_score) is swept once per Levenberg-Marquardt iteration, the 6x6 normal equations
// are reduced with warp shuffles + shared memory, one thread solves and updates the pose
// (refit_core.cuh).  fp64 throughout.  Work is tiny (<= a few thousand points x <= 20 iterations per
// problem); the point is that the pose never needs a host solve.
#include "common.cuh"
#include "refit_core.cuh"

namespace slamfe {
namespace {

constexpr int RF_THREADS = 128;

struct RefitParams {
    const double *T;          // (n_frames*H, 12) hypotheses
    int H;
    const int32_t *best;      // (n_frames, 2) [hypothesis index or -1, inlier count]
    const double *pts, *l_pix;
    const uint8_t *mask;
    const int32_t *pt_off, *pt_cnt;
    int n_points;
    double K[9];
    int max_iter;
    double tol;
    double *T_out;            // (n_frames, 12)
    int32_t *status;          // (n_frames,)
    double *rms;              // (n_frames,) or null
};

__global__ void __launch_bounds__(RF_THREADS) pnp_refit_kernel(const RefitParams p)
{
    __shared__ double s_T[12];
    __shared__ double s_red[RF_THREADS / 32][REFIT_NACC];
    __shared__ double s_acc[REFIT_NACC];
    __shared__ RefitState s_state;
    __shared__ int s_flag;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = p.pt_off ? p.pt_off[f] : 0;
    const int np = p.pt_cnt ? p.pt_cnt[f] : (p.pt_off ? p.pt_off[f + 1] - p0 : p.n_points);
    const int bi = p.best[2 * f], cnt = p.best[2 * f + 1];
    if (bi < 0 || cnt < 4 || np < 4) {  // ransac.py:187-188: fewer than 4 inliers -> no pose
        if (tid < 12) p.T_out[12 * static_cast<size_t>(f) + tid] = 0.0;
        if (tid == 0) {
            p.status[f] = 0;
            if (p.rms) p.rms[f] = 0.0;
        }
        return;
    }
    if (tid < 12) s_T[tid] = p.T[(static_cast<size_t>(f) * p.H + bi) * 12 + tid];
    if (tid == 0) {
        s_state.lambda = 1e-4;
        s_state.have_acc = 0;
        s_flag = 0;
    }
    __syncthreads();
    int it = 0, flag = 0, n_used = 0;
    for (; it < p.max_iter; ++it) {
        double acc[REFIT_NACC];
#pragma unroll
        for (int k = 0; k < REFIT_NACC; ++k) acc[k] = 0.0;
        n_used = 0;
        for (int k = tid; k < np; k += RF_THREADS) {
            const size_t g = static_cast<size_t>(p0 + k);
            if (!p.mask[g]) continue;
            refit_accumulate(s_T, p.K, p.pts[3 * g], p.pts[3 * g + 1], p.pts[3 * g + 2], p.l_pix[2 * g],
                             p.l_pix[2 * g + 1], acc);
            ++n_used;
        }
#pragma unroll
        for (int k = 0; k < REFIT_NACC; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) s_red[warp][k] = v;
        }
        __syncthreads();
        if (tid < REFIT_NACC) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < RF_THREADS / 32; ++w) v += s_red[w][tid];
            s_acc[tid] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double T_try[12];
            const int r = refit_step(s_state, s_T, s_acc, T_try, p.tol);
            if (r == 0)
                for (int k = 0; k < 12; ++k) s_T[k] = T_try[k];
            s_flag = r;
        }
        __syncthreads();
        flag = s_flag;
        if (flag != 0) break;
    }
    // the answer is the last ACCEPTED pose (the candidate of an unfinished iteration is discarded)
    if (tid < 12) p.T_out[12 * static_cast<size_t>(f) + tid] = s_state.T_acc[tid];
    if (tid == 0) {
        p.status[f] = flag < 0 ? -1 : (flag == 1 ? it + 1 : -(p.max_iter + 1));
        if (p.rms) p.rms[f] = sqrt(s_state.acc_acc[27] / fmax(1.0, static_cast<double>(cnt)));
    }
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_pnp_refit(const double *T, int H, const int32_t *best, const double *pts, const double *l_pix,
                                const uint8_t *mask, const int32_t *pt_off, const int32_t *pt_cnt, int n_points,
                                int n_frames, const double *K, int max_iter, double tol, double *T_out,
                                int32_t *status, double *rms, slamfe_stream_t stream)
{
    if (H < 0 || n_frames < 0 || n_points < 0 || max_iter < 1 || !(tol >= 0.0)) return SLAMFE_EINVAL;
    if (n_frames == 0) return 0;
    if (!T || !best || !K || !T_out || !status) return SLAMFE_EINVAL;
    if (!pt_off && n_frames != 1) return SLAMFE_EINVAL;
    if (pt_cnt && !pt_off) return SLAMFE_EINVAL;
    if (!pts || !l_pix || !mask) return SLAMFE_EINVAL;
    RefitParams p{};
    p.T = T; p.H = H; p.best = best; p.pts = pts; p.l_pix = l_pix; p.mask = mask;
    p.pt_off = pt_off; p.pt_cnt = pt_cnt; p.n_points = n_points;
    for (int k = 0; k < 9; ++k) p.K[k] = K[k];
    p.max_iter = max_iter; p.tol = tol; p.T_out = T_out; p.status = status; p.rms = rms;
    pnp_refit_kernel<<<n_frames, RF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return launch_status();
}
