// RANSAC-PnP hypothesis generation on the GPU — sm_100a.
//
// The host half of the reference's RANSAC loops (final_project/algorithms/ransac.py:94-104 and
// :155-171: np.random.choice(n, 4) + cv2.solvePnP(EPNP) + rodriguez_to_mat, 128 us per hypothesis
// on the CPU) as one launch: one thread per (frame, hypothesis) draws 4 distinct correspondences
// with a counter-based RNG (or takes them from `sample_idx`), solves P3P on the first three and
// lets the fourth pick the pose (csrc/p3p.cuh, fp64 registers).  The output feeds
// slamfe_ransac_score directly, so a whole RANSAC-PnP stays on the device.
// Not bit-comparable with OpenCV's EPnP (see p3p.cuh); contract in DESIGN.md section 2.6.
#include "common.cuh"
#include "p3p.cuh"

namespace slamfe {
namespace {

struct GenParams {
    const double *pts, *l_pix;
    const int32_t *pt_off, *pt_cnt;
    int n_points, H;
    const int32_t *sample_idx;
    const int32_t *n_hyp;  // per-frame number of hypotheses to generate (<= H), or null = H
    unsigned long long seed;
    int frame_base;  // global index of frame 0 of this launch (RNG stream id)
    double K[9], Kinv[9];
    double *T;
    uint8_t *valid;
};

constexpr int GEN_THREADS = 128;

__global__ void __launch_bounds__(GEN_THREADS) ransac_hypotheses_kernel(const GenParams p)
{
    const int f = blockIdx.y;
    const int h = blockIdx.x * GEN_THREADS + threadIdx.x;
    if (h >= p.H) return;
    const int p0 = p.pt_off ? p.pt_off[f] : 0;
    const int np = p.pt_cnt ? p.pt_cnt[f] : (p.pt_off ? p.pt_off[f + 1] - p0 : p.n_points);
    const size_t o = static_cast<size_t>(f) * p.H + h;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = 0.0;
    bool ok = np >= 4 && (!p.n_hyp || h < p.n_hyp[f]);
    if (ok) {
        int idx[4];
        if (p.sample_idx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) idx[k] = p.sample_idx[4 * o + k];
#pragma unroll
            for (int k = 0; k < 4; ++k) ok = ok && idx[k] >= 0 && idx[k] < np;
        } else {
            p3p::sample4(p.seed, static_cast<uint32_t>(p.frame_base + f), static_cast<uint32_t>(h), np, idx);
        }
        if (ok) {
            p3p::Vec3 P[4];
            double uv[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const size_t g = static_cast<size_t>(p0 + idx[k]);
                P[k] = {p.pts[3 * g], p.pts[3 * g + 1], p.pts[3 * g + 2]};
                uv[k][0] = p.l_pix[2 * g];
                uv[k][1] = p.l_pix[2 * g + 1];
            }
            ok = p3p::solve_sample(P, uv, p.K, p.Kinv, T);
            if (!ok) {
#pragma unroll
                for (int k = 0; k < 12; ++k) T[k] = 0.0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) p.T[12 * o + k] = T[k];
    p.valid[o] = ok ? 1 : 0;
}

bool invert3(const double *m, double *inv)
{
    const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
    if (det == 0.0 || det != det) return false;
    const double s = 1.0 / det;
    inv[0] = c0 * s; inv[1] = (m[2] * m[7] - m[1] * m[8]) * s; inv[2] = (m[1] * m[5] - m[2] * m[4]) * s;
    inv[3] = c1 * s; inv[4] = (m[0] * m[8] - m[2] * m[6]) * s; inv[5] = (m[2] * m[3] - m[0] * m[5]) * s;
    inv[6] = c2 * s; inv[7] = (m[1] * m[6] - m[0] * m[7]) * s; inv[8] = (m[0] * m[4] - m[1] * m[3]) * s;
    return true;
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_ransac_hypotheses(const double *pts, const double *l_pix, const int32_t *pt_off,
                                        const int32_t *pt_cnt, int n_points, int n_frames, int H,
                                        const int32_t *n_hyp, const int32_t *sample_idx, uint64_t seed,
                                        int frame_index_base, const double *K, double *T, uint8_t *hyp_valid,
                                        slamfe_stream_t stream)
{
    if (H < 0 || n_frames < 0 || n_points < 0) return SLAMFE_EINVAL;
    if (H == 0 || n_frames == 0) return 0;
    if (!pts || !l_pix || !K || !T || !hyp_valid) return SLAMFE_EINVAL;
    if (!pt_off && n_frames != 1) return SLAMFE_EINVAL;
    if (pt_cnt && !pt_off) return SLAMFE_EINVAL;
    GenParams p{};
    p.pts = pts; p.l_pix = l_pix; p.pt_off = pt_off; p.pt_cnt = pt_cnt; p.n_points = n_points; p.H = H;
    p.sample_idx = sample_idx; p.n_hyp = n_hyp; p.seed = seed; p.frame_base = frame_index_base; p.T = T;
    p.valid = hyp_valid;
    for (int k = 0; k < 9; ++k) p.K[k] = K[k];
    if (!invert3(K, p.Kinv)) return SLAMFE_EINVAL;
    const dim3 grid((H + GEN_THREADS - 1) / GEN_THREADS, n_frames);
    if (grid.y > 65535u) return SLAMFE_ERANGE;
    ransac_hypotheses_kernel<<<grid, GEN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return launch_status();
}
