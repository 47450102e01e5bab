// Loop-closure candidate gating on the device — the arithmetic of get_good_candidates / check_candidate
// (final_project/backend/loop/loop_closure.py:164-228) for MANY query keyframes in one launch.
//
// The reference, per query keyframe c_n and per earlier keyframe c_i (i < n - KEY_FRAME_GAP):
//   * shortest path c_i -> c_n in the covariance graph (backend/loop/graph.py:57-97: Dijkstra with edge
//     weight det(cov), heap order (distance, node), strict-< relaxation);
//   * relative covariance = sum of the edge covariances along that path, in path order
//     (get_relative_covariance_along_path, loop_closure.py:103-137);
//   * Mahalanobis distance sqrt(2 * BetweenFactorPose3(c_n, c_i, Pose3(), Gaussian(cov)).error(result))
//     (:185-188) = sqrt(xi^T cov^-1 xi), xi = Pose3::Logmap(pose_n^-1 pose_i) (GTSAM's full exponential
//     chart, tangent order (rotation, translation)).
// The GTSAM objects (optimised poses, marginal covariances of consecutive keyframes) stay with the caller
// and arrive as arrays.  One CTA per query: Dijkstra FROM c_n over <= 1024 nodes (in an undirected graph
// with positive weights the tree from c_n holds the same paths), then one thread per candidate walks its
// path towards c_n summing covariances in the reference's order and solves the 6x6 system.  fp64.
#include "common.cuh"

namespace slamfe {
namespace {

constexpr int GT_THREADS = 256;
constexpr int GT_MAX_NODES = 1024;

struct GateParams {
    const double *poses;       // (K, 12) camera-to-world [R|t] of every keyframe (result.atPose3)
    const int32_t *adj_off;    // (K + 1,) CSR of the undirected covariance graph
    const int32_t *adj_node;   // (2E,) neighbour
    const int32_t *adj_edge;   // (2E,) edge id
    const double *edge_w;      // (E,) det(cov)
    const double *edge_cov;    // (E, 36)
    const int32_t *queries;    // (Q,) query keyframe index n
    int n_nodes, gap;
    double *dist_out;          // (Q, K) Mahalanobis distance, +inf where i is not a candidate
    int32_t *hops_out;         // (Q, K) path length in edges (diagnostic), -1 where not a candidate
};

// SO(3) / SE(3) logarithms as GTSAM computes them (SO3::Logmap, Pose3::Logmap).
__device__ void pose_logmap(const double *R, const double *t, double *xi)
{
    const double tr = R[0] + R[4] + R[8];
    double w[3];
    if (tr + 1.0 < 1e-3) {  // rotation close to pi: use the largest diagonal element
        int k = 0;
        if (R[4] > R[0]) k = 1;
        if (R[8] > R[4 * k]) k = 2;
        const double d = R[4 * k];
        const double s = 3.14159265358979323846 / sqrt(2.0 + 2.0 * d);
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        w[k] = s * (1.0 + d);
        w[a] = s * R[3 * a + k];
        w[b] = s * R[3 * b + k];
        // sign fix: make the log consistent with the skew part where it is non-zero
        const double sk = (k == 0) ? R[7] - R[5] : (k == 1) ? R[2] - R[6] : R[3] - R[1];
        if (sk < 0) { w[0] = -w[0]; w[1] = -w[1]; w[2] = -w[2]; }
    } else {
        double mag;
        const double tr3 = tr - 3.0;
        if (tr3 < -1e-7) {
            const double theta = acos((tr - 1.0) / 2.0);
            mag = theta / (2.0 * sin(theta));
        } else {
            mag = 0.5 - tr3 / 12.0;   // Taylor expansion near the identity
        }
        w[0] = mag * (R[7] - R[5]);
        w[1] = mag * (R[2] - R[6]);
        w[2] = mag * (R[3] - R[1]);
    }
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
    if (th < 1e-10) {
        xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2];
        return;
    }
    // u = T - (th/2) W T + (1 - th / (2 tan(th/2))) W W T,  W = skew(w / th)
    const double n0 = w[0] / th, n1 = w[1] / th, n2 = w[2] / th;
    const double wt0 = n1 * t[2] - n2 * t[1], wt1 = n2 * t[0] - n0 * t[2], wt2 = n0 * t[1] - n1 * t[0];
    const double wwt0 = n1 * wt2 - n2 * wt1, wwt1 = n2 * wt0 - n0 * wt2, wwt2 = n0 * wt1 - n1 * wt0;
    const double c = 1.0 - th / (2.0 * tan(0.5 * th));
    xi[3] = t[0] - 0.5 * th * wt0 + c * wwt0;
    xi[4] = t[1] - 0.5 * th * wt1 + c * wwt1;
    xi[5] = t[2] - 0.5 * th * wt2 + c * wwt2;
}

// xi^T S^-1 xi by Cholesky; NaN when S is not positive definite
__device__ double mahalanobis2(const double *S, const double *xi)
{
    double L[6][6];
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = S[6 * r + c];
            for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
            if (r == c) {
                if (!(s > 0.0)) return __longlong_as_double(0x7FF8000000000000ll);
                L[r][r] = sqrt(s);
            } else {
                L[r][c] = s / L[c][c];
            }
        }
    double y[6], acc = 0.0;
    for (int r = 0; r < 6; ++r) {
        double s = xi[r];
        for (int k = 0; k < r; ++k) s -= L[r][k] * y[k];
        y[r] = s / L[r][r];
        acc += y[r] * y[r];
    }
    return acc;
}

__global__ void __launch_bounds__(GT_THREADS) gate_kernel(const GateParams p)
{
    __shared__ double s_dist[GT_MAX_NODES];
    __shared__ int s_pred[GT_MAX_NODES], s_pred_edge[GT_MAX_NODES];
    __shared__ unsigned char s_done[GT_MAX_NODES];
    __shared__ double s_red_d[GT_THREADS / 32];
    __shared__ int s_red_n[GT_THREADS / 32];
    __shared__ int s_cur;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.n_nodes, n = p.queries[q];
    const double INF = __longlong_as_double(0x7FF0000000000000ll);
    for (int v = tid; v < K; v += GT_THREADS) {
        s_dist[v] = v == n ? 0.0 : INF;
        s_pred[v] = -1;
        s_pred_edge[v] = -1;
        s_done[v] = 0;
    }
    __syncthreads();
    // ---- Dijkstra from n: finalise nodes in (distance, node) order, strict-< relaxation (graph.py:72-85) ----
    for (int it = 0; it < K; ++it) {
        double bd = INF;
        int bn = 0x7FFFFFFF;
        for (int v = tid; v < K; v += GT_THREADS)
            if (!s_done[v] && (s_dist[v] < bd || (s_dist[v] == bd && v < bn))) { bd = s_dist[v]; bn = v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xFFFFFFFFu, bd, o);
            const int on = __shfl_xor_sync(0xFFFFFFFFu, bn, o);
            if (od < bd || (od == bd && on < bn)) { bd = od; bn = on; }
        }
        if (lane == 0) { s_red_d[warp] = bd; s_red_n[warp] = bn; }
        __syncthreads();
        if (tid == 0) {
            double d = s_red_d[0];
            int v = s_red_n[0];
            for (int w = 1; w < GT_THREADS / 32; ++w)
                if (s_red_d[w] < d || (s_red_d[w] == d && s_red_n[w] < v)) { d = s_red_d[w]; v = s_red_n[w]; }
            s_cur = (d < INF) ? v : -1;
            if (s_cur >= 0) s_done[s_cur] = 1;
        }
        __syncthreads();
        const int u = s_cur;
        if (u < 0) break;  // the rest is unreachable
        const double du = s_dist[u];
        for (int e = p.adj_off[u] + tid; e < p.adj_off[u + 1]; e += GT_THREADS) {
            const int v = p.adj_node[e];
            const double nd = du + p.edge_w[p.adj_edge[e]];
            if (!s_done[v] && nd < s_dist[v]) {   // neighbours of u are distinct: no write conflicts
                s_dist[v] = nd;
                s_pred[v] = u;
                s_pred_edge[v] = p.adj_edge[e];
            }
        }
        __syncthreads();
    }
    // ---- candidates: path i -> ... -> n, covariances summed in path order, Mahalanobis distance ----
    const double *Pn = p.poses + 12 * static_cast<size_t>(n);
    for (int i = tid; i < K; i += GT_THREADS) {
        double out = INF;
        int hops = -1;
        if (i < n - p.gap && s_dist[i] < INF) {
            double S[36];
            hops = 0;
            for (int v = i; v != n; v = s_pred[v]) {
                const double *c = p.edge_cov + 36 * static_cast<size_t>(s_pred_edge[v]);
                if (hops == 0) {
#pragma unroll
                    for (int k = 0; k < 36; ++k) S[k] = c[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 36; ++k) S[k] = S[k] + c[k];
                }
                ++hops;
            }
            // relative pose pose_n.between(pose_i) = [Rn^T Ri | Rn^T (ti - tn)]
            const double *Pi = p.poses + 12 * static_cast<size_t>(i);
            double R[9], t[3], xi[6];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c)
                    R[3 * r + c] = Pn[r] * Pi[c] + Pn[4 + r] * Pi[4 + c] + Pn[8 + r] * Pi[8 + c];
                t[r] = Pn[r] * (Pi[3] - Pn[3]) + Pn[4 + r] * (Pi[7] - Pn[7]) + Pn[8 + r] * (Pi[11] - Pn[11]);
            }
            pose_logmap(R, t, xi);
            out = sqrt(mahalanobis2(S, xi));
        }
        p.dist_out[static_cast<size_t>(q) * K + i] = out;
        if (p.hops_out) p.hops_out[static_cast<size_t>(q) * K + i] = hops;
    }
}

}  // namespace
}  // namespace slamfe

using namespace slamfe;

extern "C" int slamfe_gate_candidates(const double *poses, int n_nodes, const int32_t *adj_off, const int32_t *adj_node,
                                      const int32_t *adj_edge, const double *edge_w, const double *edge_cov,
                                      const int32_t *queries, int n_queries, int gap, double *dist_out,
                                      int32_t *hops_out, slamfe_stream_t stream)
{
    if (n_nodes < 0 || n_queries < 0 || gap < 0) return SLAMFE_EINVAL;
    if (n_queries == 0 || n_nodes == 0) return 0;
    if (n_nodes > GT_MAX_NODES) return SLAMFE_ERANGE;
    if (!poses || !adj_off || !adj_node || !adj_edge || !edge_w || !edge_cov || !queries || !dist_out)
        return SLAMFE_EINVAL;
    GateParams p{};
    p.poses = poses; p.adj_off = adj_off; p.adj_node = adj_node; p.adj_edge = adj_edge; p.edge_w = edge_w;
    p.edge_cov = edge_cov; p.queries = queries; p.n_nodes = n_nodes; p.gap = gap; p.dist_out = dist_out;
    p.hops_out = hops_out;
    gate_kernel<<<n_queries, GT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return launch_status();
}
