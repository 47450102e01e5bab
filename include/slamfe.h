/*
 * slamfe.h — C-ABI of libslamfe.so, the B200 (sm_100a) front-end kernels.
 *
 * This is the drop-in boundary for the reference's front-end hot path
 * (michaelpiro/67604-SLAM---video-navigation, final_project/algorithms/*).  The reference has no
 * native FFI: its hot path sits behind Python module attributes, and the
 * arithmetic runs inside OpenCV / NumPy.  Each entry point below names the reference call it
 * replaces; the Python mirror in `67604-slam---video-navigation_b200/` binds them through ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every data pointer is a DEVICE pointer owned by the caller (who keeps it alive until the
 *     stream has drained), except small camera matrices, which are HOST pointers read at call
 *     time (documented per function).
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *   - Return value: 0 on success, a negative SLAMFE_E* code for argument errors, or a positive
 *     cudaError_t.  Nothing throws; no persistent allocations; no global mutable state.
 *
 * Packed match keys
 *   The matcher's native result is a 32-bit key per neighbour:
 *       key = (hamming_distance << 22) | train_index        (distance <= 512, index < 2^22)
 *       SLAMFE_KEY_NONE (0xFFFFFFFF) = no neighbour.
 *   Unsigned comparison of keys == cv2.BFMatcher ordering: smaller distance first, ties broken
 *   by the LOWER train index (first minimum).  Min-merging keys across train shards / GPUs is
 *   therefore exact.
 */
#ifndef SLAMFE_H
#define SLAMFE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *slamfe_stream_t;

#define SLAMFE_KEY_IDX_BITS 22
#define SLAMFE_KEY_IDX_MASK 0x3FFFFFu
#define SLAMFE_KEY_NONE 0xFFFFFFFFu
#define SLAMFE_MAX_DESC_BYTES 64

/* matcher flags */
#define SLAMFE_MATCH_BEST_ONLY 1 /* keep only the best neighbour (.match / crossCheck); row_keys[:,1] = KEY_NONE */
#define SLAMFE_MATCH_COMPACT_KEYS 2 /* with BEST_ONLY: row_keys is (rows,) uint32, one key per query row */
#define SLAMFE_MATCH_MMA 4 /* run the sweep on the tcgen05 tensor cores (int8 contraction, identical keys) */

#define SLAMFE_EINVAL (-1)   /* bad argument (null pointer, negative size, stride < desc_bytes ...) */
#define SLAMFE_ERANGE (-2)   /* size exceeds what the key encoding / grid can address */

/* ABI version of this header: bumped whenever an entry point is added or a signature changes.  The loader
 * (_cabi.load_library) refuses a library whose slamfe_version() differs. */
#define SLAMFE_ABI_VERSION 205
int slamfe_version(void);
const char *slamfe_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Hamming matcher
 * ---------------------------------------------------------------------------------------- */

/*
 * Brute-force Hamming top-2 of every query row against every train row.
 * Replaces cv2.BFMatcher(NORM_HAMMING).match / .knnMatch(k=2):
 *   final_project/backend/database/database.py:54-55, backend/loop/loop_closure.py:422,
 *   final_project/algorithms/matching.py:15,44, VAN_ex/code/ex1.py:189-190.
 *
 *   q, t          descriptor rows, `desc_bytes` (<= 64) useful bytes every `*_stride` bytes
 *                 (cv2 layout: stride 61 for AKAZE MLDB; 64-byte padded rows also accepted).
 *   t_index_base  added to the train index stored in the keys (global index of t[0] when the
 *                 train set is one shard of a larger one).
 *   row_keys      out, (nq, 2) uint32: best and second-best key per query row.
 *   col_keys      out or NULL, (nt,) uint32: per TRAIN row the best (distance<<22 | query_index),
 *                 i.e. the result of matching t against q, from the same single pass.  This is
 *                 what crossCheck (matching.py:22,44) and the backward match (database.py:55)
 *                 need; cv2 spends a second full pass on it.
 *   flags         0, or SLAMFE_MATCH_BEST_ONLY when the second neighbour is not needed
 *                 (.match() and crossCheck; knnMatch(k=2) / the ratio test need 0).
 */
int slamfe_hamming_top2(const uint8_t *q, int nq, int q_stride,
                        const uint8_t *t, int nt, int t_stride,
                        int desc_bytes, int t_index_base,
                        uint32_t *row_keys, uint32_t *col_keys, int flags, slamfe_stream_t stream);

/*
 * Ragged batch of independent (query set, train set) problems in ONE launch — one problem per
 * frame pair / loop-closure candidate pair.  Problem p matches rows
 * [q_off[p], q_off[p] + nq_p) of `q` against rows [t_off[p], t_off[p] + nt_p) of `t`, where
 * nq_p = q_cnt ? q_cnt[p] : q_off[p+1] - q_off[p]  (same for t).  All four arrays are DEVICE
 * int32; q_off/t_off have n_problems + 1 entries unless the matching *_cnt array is given, in which
 * case n_problems entries suffice.  Keys hold problem-local train indices + t_index_base (the global
 * index of every problem's first train row when the train sets are one shard of larger ones; same
 * base for all problems).  row_keys is
 * (q_rows_total, 2), col_keys (t_rows_total,) or NULL; both are indexed by global row.
 * max_nq / max_nt are host-side upper bounds on any problem's size (they size the grid).
 */
int slamfe_hamming_top2_batched(const uint8_t *q, int q_stride, const int32_t *q_off, const int32_t *q_cnt,
                                const uint8_t *t, int t_stride, const int32_t *t_off, const int32_t *t_cnt,
                                int n_problems, int max_nq, int max_nt, int desc_bytes, int t_index_base,
                                uint32_t *row_keys, int64_t q_rows_total,
                                uint32_t *col_keys, int64_t t_rows_total, int flags, slamfe_stream_t stream);

/*
 * Candidate-pair form (loop closure, backend/loop/loop_closure.py:422 inside :405-436, and BASELINE
 * config 4 "each keyframe vs all prior keyframes"): problem p matches q rows
 * [q_off[p], q_off[p] + q_cnt[p]) against t rows [t_off[p], t_off[p] + t_cnt[p]); the same rows may
 * appear in many problems (q and t may be the same keyframe pool), so the result rows are indexed by
 * PROBLEM: the keys of problem p's query i land in row_keys[out_off[p] + i] ((out_rows_total, 2)).
 * All five index arrays are DEVICE int32 with n_problems entries.  No column minima
 * (MATCHER.match at loop_closure.py:422 has crossCheck=False).
 */
int slamfe_hamming_top2_pairs(const uint8_t *q, int q_stride, const int32_t *q_off, const int32_t *q_cnt,
                              const uint8_t *t, int t_stride, const int32_t *t_off, const int32_t *t_cnt,
                              const int32_t *out_off, int n_problems, int max_nq, int max_nt, int desc_bytes,
                              uint32_t *row_keys, int64_t out_rows_total, int flags, slamfe_stream_t stream);

/* keys (n,) -> idx (n,) int32 (-1 for NONE), dist (n,) int32 (-1 for NONE). */
int slamfe_unpack_keys(const uint32_t *keys, int64_t n, int32_t *idx, int32_t *dist, slamfe_stream_t stream);

/*
 * Merge per-shard top-2 tables after an all-gather: shard_keys is (n_shards, nq, 2); out (nq, 2)
 * receives the two smallest keys per query over all shards (exact: keys carry global indices).
 */
int slamfe_merge_top2(const uint32_t *shard_keys, int n_shards, int nq, uint32_t *out, slamfe_stream_t stream);

/*
 * crossCheck epilogue (cv2.BFMatcher(crossCheck=True), matching.py:22,44; and the manual
 * forward/backward loop database.py:67-77): match_t[i] = train index of query i if the pair is
 * mutual (col_keys[best_i] points back at i) else -1; match_dist[i] likewise.
 */
int slamfe_cross_check(const uint32_t *row_keys, const uint32_t *col_keys, int nq, int nt,
                       int32_t *match_t, int32_t *match_dist, slamfe_stream_t stream);

/* Ratio test of VAN_ex/code/ex1.py:118-122: mask[i] = (num * d1 < den * d2), e.g. 5*d1 < 3*d2 for 0.6. */
int slamfe_ratio_test(const uint32_t *row_keys, int nq, int num, int den, uint8_t *mask, slamfe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rectified-stereo row filter and links
 * ---------------------------------------------------------------------------------------- */

/*
 * extract_inliers_outliers (final_project/algorithms/matching.py:48-69): for match m between
 * left keypoint match_q[m] and right keypoint match_t[m]:
 *   mask[m] = |yl - yr| < 2  &&  xl > xr + 2        (matching.py:62-63)
 * pts are (n, 2) float32 (x, y) — cv2.KeyPoint.pt values.
 */
int slamfe_stereo_filter(const float *pts_left, const float *pts_right,
                         const int32_t *match_q, const int32_t *match_t, int n_matches,
                         uint8_t *mask, slamfe_stream_t stream);

/*
 * Fused per-frame epilogue of the stereo match, batched over frames (one CTA per frame):
 * crossCheck (matching.py:44) + row filter (matching.py:48-69) + TrackingDB.create_links
 * (backend/database/tracking_database.py:224-246) + features[is_valid] compaction.
 * Inputs are the keys of slamfe_hamming_top2_batched(left, right) with the same row offsets
 * l_off / r_off and optional per-frame counts l_cnt / r_cnt (NULL: count = offset difference).
 * Outputs, all indexed with l_off as per-frame capacity (rows past n_links[f] inside a frame's
 * capacity are filled with link_src = -1, links = 0):
 *   match_t   (l_rows,)   int32  mutual right index per left row, -1 if not mutual
 *   n_matches (n_frames,) int32  number of mutual matches      (len(matches), database.py:26)
 *   n_links   (n_frames,) int32  number of stereo inliers
 *   link_src  (l_rows,)   int32  left keypoint index of the k-th link, ascending
 *   links     (l_rows, 3) float32 [x_left, x_right, (yl + yr) / 2]  (tracking_database.py:243)
 *   feat      (l_rows, 64) uint8 or NULL: compacted descriptors, zero-padded to 64-byte rows
 */
int slamfe_stereo_links_batched(const uint32_t *row_keys, const uint32_t *col_keys,
                                const int32_t *l_off, const int32_t *l_cnt,
                                const int32_t *r_off, const int32_t *r_cnt, int n_frames,
                                const float *pts_left, const float *pts_right,
                                const uint8_t *desc_left, int l_stride, int desc_bytes,
                                int32_t *match_t, int32_t *n_matches, int32_t *n_links,
                                int32_t *link_src, float *links, uint8_t *feat, slamfe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Triangulation   (P, Q are HOST pointers to 12 doubles, row-major 3x4)
 * ---------------------------------------------------------------------------------------- */

/*
 * triangulate_links / triangulate_last_frame (final_project/algorithms/triangulation.py:27-50):
 * DLT of (x_left, y), (x_right, y).  For links rows 1 and 3 of the DLT matrix coincide
 * (P[1]==Q[1], P[2]==Q[2]), so the 4x4 null vector is the exact cofactor vector of the three
 * distinct rows; computed in fp64 registers.  Returns SLAMFE_EINVAL if P[1:]!=Q[1:] (use the
 * general entry point).  links: (n, 3) [x_left, x_right, y]; xyz: (n, 3).
 */
int slamfe_triangulate_links_f64(const double *links, int64_t n, const double *P, const double *Q,
                                 double *xyz, slamfe_stream_t stream);
int slamfe_triangulate_links_f32(const float *links, int64_t n, const double *P, const double *Q,
                                 float *xyz, slamfe_stream_t stream);

/*
 * linear_least_squares_triangulation (triangulation.py:5-24) for arbitrary P, Q and distinct
 * y (analysis.py:400, VAN_ex/code/ex2.py:209): smallest right singular vector of the 4x4 DLT
 * matrix by one-sided Jacobi in fp64 registers.  pxy, qxy: (n, 2); xyz: (n, 3).
 */
int slamfe_triangulate_dlt_f64(const double *pxy, const double *qxy, int64_t n,
                               const double *P, const double *Q, double *xyz, slamfe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * RANSAC-PnP hypothesis scoring   (K: 9, M1/M2: 12 HOST doubles, row-major)
 * ---------------------------------------------------------------------------------------- */

/*
 * transformation_agreement x all hypotheses (final_project/algorithms/ransac.py:28-56 inside the
 * loops ransac.py:94-112 and :155-182), batched over `n_frames` independent problems in ONE
 * launch.  Frame f owns hypotheses T[f*H .. f*H+H) (each 3x4 row-major fp64), and points
 * [pt_off[f], pt_off[f+1]) of pts (.,3) / l_pix (.,2) / r_pix (.,2) (fp64).  pt_off is a DEVICE
 * int32 array of n_frames + 1 entries, or NULL when n_frames == 1 (then n_points is used); with
 * pt_cnt (n_frames,) the frame's point count is pt_cnt[f] instead of the offset difference (ragged
 * tables with per-frame capacity, as slamfe_track_gather writes them; pt_off then needs n_frames entries).
 *   hyp_valid  (n_frames*H,) uint8 or NULL: 0 = hypothesis skipped (solvePnP failed, ransac.py:101-104)
 *   counts     out (n_frames*H,) int32: inlier count per hypothesis (np.sum, ransac.py:109)
 *   best       out (n_frames, 2) int32: [index of the first hypothesis with the largest count
 *              (ransac.py:110 keeps strictly-better only), that count]; index -1 if no
 *              hypothesis scored > 0 inliers
 *   best_mask  out (total_points,) uint8: inlier mask of the best hypothesis (ransac.py:112)
 *   work       scratch, (n_frames,) int32, zeroed by the call
 * Every counted verdict is the reference's: fp64, association order ((K@T)@M_h)@X, IEEE division, strict < 2
 * (ransac.py:38-56), evaluated division-free with a rounding certificate and the literal division as fallback;
 * pairs that a one-sided fp32 pre-filter proves outside never reach the fp64 path (csrc/ransac_core.cuh).
 */
int slamfe_ransac_score(const double *T, const uint8_t *hyp_valid, int H,
                        const double *pts, const double *l_pix, const double *r_pix,
                        const int32_t *pt_off, const int32_t *pt_cnt, int n_points, int n_frames, int max_points,
                        const double *K, const double *M1, const double *M2,
                        int32_t *counts, int32_t *best, uint8_t *best_mask, int32_t *work,
                        slamfe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Frame-to-frame tracking glue (the create_db loop body, backend/database/database.py:48-87)
 * ---------------------------------------------------------------------------------------- */

/*
 * For every consecutive frame pair (f, f+1), f < n_pairs, from the tables of the sequence pipeline
 * (slamfe_hamming_top2_batched on the compacted features + slamfe_stereo_links_batched):
 *   - mutual forward/backward check (database.py:67-77): forward match j kept iff
 *     bwd[fwd[j]] == j; kept matches in ascending j (== good_idx);
 *   - link gather + pixel arrays (ransac.py:76-81, :85-88) and fp64 triangulation of the previous
 *     frame's links (ransac.py:83) from the exact Link values (x float32, y = (yl+yr)/2 in double);
 *   - n_hyp[f] = min(h_max, calc_ransac_iteration(100 * n_links[f+1] / n_matches[f+1]))
 *     (ransac.py:59-67, database.py:26,80); n_hyp_full[f] (optional, may be NULL) receives the
 *     UNCAPPED iteration count, so the caller can see which pairs h_max truncated (n_hyp_full[f] >
 *     h_max) and re-run those at their full count — the reference always runs the full count.
 * fwd_keys (rows,) uint32: COMPACT best keys of the forward pass (SLAMFE_MATCH_BEST_ONLY | SLAMFE_MATCH_COMPACT_KEYS),
 * bwd_keys (rows,) uint32: its column minima.
 * Row-indexed outputs use frame f's row offset l_off[f] as base and hold n_good[f] entries:
 *   good_j, good_t (L,) int32 link indices in frame f / f+1;  pts (L,3), lpix (L,2), rpix (L,2) fp64.
 * They are the pts / l_pix / r_pix (pt_off = l_off, pt_cnt = n_good) of slamfe_ransac_hypotheses and
 * slamfe_ransac_score.  P, Q: HOST 3x4 (shared rows 1, 2 as for slamfe_triangulate_links_*).
 */
int slamfe_track_gather(const uint32_t *fwd_keys, const uint32_t *bwd_keys, const int32_t *l_off, const int32_t *r_off,
                        const int32_t *n_links, const int32_t *n_matches, const float *pts_left,
                        const float *pts_right, const int32_t *link_src, const int32_t *match_t, int n_pairs,
                        const double *P, const double *Q, int h_max, int32_t *good_j, int32_t *good_t,
                        int32_t *n_good, int32_t *n_hyp, int32_t *n_hyp_full, double *pts, double *lpix,
                        double *rpix, slamfe_stream_t stream);

/*
 * Loop-closure candidate gather (check_candidate_match, backend/loop/loop_closure.py:405-436, feeding
 * ransac_pnp, ransac.py:132-146): candidate p = (keyframe A at rows q_off[p], q_cnt[p] links; keyframe B
 * at rows t_off[p]); keys = compact best keys of slamfe_hamming_top2_pairs (one per query, at
 * out_off[p] + i).  kf_pts (.,3): triangulated links of every keyframe (slamfe_triangulate_links_f64),
 * kf_links (.,3) fp64 [x_left, x_right, y].  Writes pts / lpix / rpix rows at out_off[p] + i:
 * A's 3-D point i and B's link pixels at trainIdx — the inputs of slamfe_ransac_hypotheses /
 * slamfe_ransac_score with pt_off = out_off, pt_cnt = q_cnt.
 */
int slamfe_pairs_gather(const uint32_t *keys, const int32_t *q_off, const int32_t *q_cnt, const int32_t *t_off,
                        const int32_t *out_off, int n_problems, int max_nq, const double *kf_pts,
                        const double *kf_links, double *pts, double *lpix, double *rpix, slamfe_stream_t stream);

/*
 * in_prev_cur of database.py:84-85: inlier_fwd[l_off[f] + good_j[k]] = 1 for the mutual matches k of
 * pair f that the best hypothesis accepts (best_mask from slamfe_ransac_score), 0 elsewhere
 * (rows_total bytes are cleared first).  When no hypothesis scored > 0 inliers (best[f][0] < 0) every
 * mutual match is flagged, as good_idx[None] does in the reference (database.py:82).
 */
int slamfe_scatter_inliers(const uint8_t *best_mask, const int32_t *good_j, const int32_t *l_off,
                           const int32_t *n_good, const int32_t *best, int n_pairs, uint8_t *inlier_fwd,
                           int64_t rows_total, slamfe_stream_t stream);

/*
 * RANSAC-PnP hypothesis generation (the host half of the loops ransac.py:94-104, :155-171:
 * np.random.choice(n, 4) + cv2.solvePnP(EPNP) on the 4 points + rodriguez_to_mat, utils.py:16-18).
 * One thread per (frame, hypothesis): 4 distinct correspondences of the frame (from `sample_idx`
 * (n_frames*H, 4) int32 if given, else drawn with a counter-based RNG from `seed`), P3P on the first
 * three, the fourth picks the pose by left-image reprojection error.  NOT bit-comparable with
 * OpenCV's EPnP (implementation-defined on 4 points; the reference samples unseeded): exact minimal
 * solver, pinned against cv2.SOLVEPNP_P3P and ground truth.
 *   pts (.,3), l_pix (.,2) fp64 DEVICE; frame f owns points [pt_off[f], pt_off[f] + n_f) with
 *   n_f = pt_cnt ? pt_cnt[f] : pt_off[f+1] - pt_off[f]  (pt_off NULL: one frame of n_points points)
 *   seed, frame_index_base: the sample of hypothesis h of frame f is a pure function of
 *            (seed, frame_index_base + f, h), so a sequence processed in chunks draws the same samples
 *   n_hyp    (n_frames,) int32 DEVICE or NULL: hypotheses wanted for frame f (<= H); the rest of the
 *            frame's H slots are marked invalid (calc_ransac_iteration differs per frame, ransac.py:59-67)
 *   T        out (n_frames*H, 12) fp64 row-major [R|t], zero where invalid
 *   hyp_valid out (n_frames*H,) uint8: 0 = degenerate sample / no admissible pose / fewer than 4 points
 * The outputs are the T / hyp_valid inputs of slamfe_ransac_score.  K: 9 HOST doubles.
 */
int slamfe_ransac_hypotheses(const double *pts, const double *l_pix, const int32_t *pt_off, const int32_t *pt_cnt,
                             int n_points, int n_frames, int H, const int32_t *n_hyp, const int32_t *sample_idx,
                             uint64_t seed, int frame_index_base, const double *K, double *T, uint8_t *hyp_valid,
                             slamfe_stream_t stream);

/*
 * Track ids of a whole sequence — the bookkeeping of TrackingDB.add_frame
 * (backend/database/tracking_database.py:273-337) over the tables of the tracking stages: link k of frame f
 * (row l_off[f] + k) continues into link j = index(fwd_keys[row]) of frame f + 1 when inlier_fwd[row] is set
 * (fwd_keys (rows_total,): COMPACT best keys, SLAMFE_MATCH_BEST_ONLY | SLAMFE_MATCH_COMPACT_KEYS)
 * (in_prev_cur, database.py:84-85; mutual matches, hence one-to-one).  A link with an inlier successor and no
 * inlier predecessor starts a track; ids are issued as issue_trackId does (:196-198): frame pairs in order,
 * ascending previous-feature index inside a pair.
 *   track_id (rows_total,) int32 out: id of the track the link belongs to, -1 (NO_ID) for links on no track
 *            and for padding rows;  n_tracks (1,) int32 out;  head_base (n_frames + 1,) int32 out: number of
 *            tracks that start before frame f
 *   pred, rank (rows_total,), head_cnt (n_frames,): int32 scratch
 */
int slamfe_track_ids(const uint32_t *fwd_keys, const uint8_t *inlier_fwd, const int32_t *l_off, const int32_t *n_links,
                     int n_frames, int64_t rows_total, int32_t *pred, int32_t *rank, int32_t *head_cnt,
                     int32_t *head_base, int32_t *track_id, int32_t *n_tracks, slamfe_stream_t stream);

/*
 * The tracking database as dense columns (what slamfe.trackdb.SoATrackingDB stores; replaces the dicts of
 * TrackingDB, tracking_database.py:75-100): the links of frame f become rows link_off[f] .. link_off[f+1] of
 *   x_left, x_right (N,) float32; y (N,) float64 = (yl + yr) / 2 (:243); feat_out (N, desc_bytes) uint8 =
 *   features[is_valid] (:235); track_out (N,) int32 (track_id of slamfe_track_ids, or -1 when track_id is NULL)
 * with link_off (n_frames + 1,) int32 = exclusive prefix sum of n_links, written by the call.  Inputs are the
 * pipeline's per-frame padded tables (link_src, match_t, feat (L, 64), keypoints).  The outputs must hold
 * sum(n_links) rows; the per-frame capacity sum (L) is always enough.
 */
int slamfe_pack_db(const int32_t *l_off, const int32_t *r_off, const int32_t *n_links, const int32_t *link_src,
                   const int32_t *match_t, const float *pts_left, const float *pts_right, const uint8_t *feat,
                   int desc_bytes, const int32_t *track_id, int n_frames, int32_t *link_off, float *x_left,
                   float *x_right, double *y, uint8_t *feat_out, int32_t *track_out, slamfe_stream_t stream);

/*
 * Loop-closure candidate gating (get_good_candidates / check_candidate, backend/loop/loop_closure.py:164-228)
 * for n_queries query keyframes in one launch.  For query n = queries[q] and every keyframe i < n - gap:
 * shortest path i -> n in the covariance graph (backend/loop/graph.py:57-97; edge weight edge_w = det(cov),
 * Dijkstra order (distance, node), strict-< relaxation), covariance = sum of edge_cov along the path in path
 * order (loop_closure.py:103-137), distance = sqrt(xi^T cov^-1 xi) with xi = Pose3::Logmap(pose_n^-1 pose_i)
 * (= sqrt(2 * BetweenFactorPose3(c_n, c_i, Pose3(), Gaussian(cov)).error(result)), :185-188; GTSAM tangent
 * order rotation, translation).  The GTSAM objects stay with the caller and arrive as arrays:
 *   poses (n_nodes, 12) fp64 camera-to-world [R|t] per keyframe;  the undirected graph as CSR over nodes
 *   (adj_off (n_nodes+1), adj_node / adj_edge (2E)) with edge_w (E,) and edge_cov (E, 36).
 *   dist_out (n_queries, n_nodes) fp64: the distance, +inf where i is no candidate (i >= n - gap, or
 *   unreachable), NaN where the summed covariance is not positive definite;  hops_out (same shape, int32 or
 *   NULL): edges on the path, -1 where no candidate.   n_nodes <= 1024.
 */
int slamfe_gate_candidates(const double *poses, int n_nodes, const int32_t *adj_off, const int32_t *adj_node,
                           const int32_t *adj_edge, const double *edge_w, const double *edge_cov,
                           const int32_t *queries, int n_queries, int gap, double *dist_out, int32_t *hops_out,
                           slamfe_stream_t stream);

/*
 * PnP refit on the consensus set: the final solve of ransac_pnp (final_project/algorithms/ransac.py:185-193,
 * cv2.solvePnP(points_3d[best_idx], l_pix[best_idx], K, EPNP)) for every frame pair / loop-closure candidate
 * of a batch in one launch, so the pose needs no host solve.  Levenberg-Marquardt on the left-image
 * reprojection error of the inliers (mask = best_mask of slamfe_ransac_score), seeded by the winning
 * hypothesis T[(f*H + best[f][0])]; one CTA per problem, fp64.  Not bit-comparable with OpenCV's EPnP
 * (a different algorithm for the same objective): tolerance contract in DESIGN.md (rotation < 1e-3 rad,
 * translation < 1 cm against cv2 on the reference's golden runs; equals an independent Gauss-Newton to 1e-9).
 *   T (n_frames*H,12), best (n_frames,2) [index or -1, inlier count], pts/l_pix/mask as for slamfe_ransac_score
 *   T_out  (n_frames,12) fp64: refit [R|t] (world -> camera, as rodriguez_to_mat, utils.py:16-18); zeros when
 *          the problem has no pose
 *   status (n_frames,) int32: k > 0 converged after k evaluations; 0 = no pose (no hypothesis or fewer than 4
 *          inliers: ransac.py:187-188 returns None); -1 = singular normal equations (T_out = the seed);
 *          -(max_iter+1) = iteration cap reached (T_out = best pose so far)
 *   rms    (n_frames,) fp64 or NULL: RMS left reprojection error of the inliers at T_out, pixels
 * K: 9 HOST doubles.  tol: relative decrease of the squared error that counts as converged (e.g. 1e-12).
 */
int slamfe_pnp_refit(const double *T, int H, const int32_t *best, const double *pts, const double *l_pix,
                     const uint8_t *mask, const int32_t *pt_off, const int32_t *pt_cnt, int n_points, int n_frames,
                     const double *K, int max_iter, double tol, double *T_out, int32_t *status, double *rms,
                     slamfe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Roofline micro-benchmarks (measure the pipe peaks the matcher / scorer are bound by)
 * ---------------------------------------------------------------------------------------- */

/* Runs `iters` dependent-chain batches of POPC (mode 0), POPC+LOP3+IADD3 matcher mix (mode 1)
 * or fp64 FMA (mode 2) on every SM; *ops_per_thread_iter (host) receives the number of
 * counted operations each thread performs per iteration.  sink: device uint32[grid*block]. */
int slamfe_peak_kernel(int mode, int iters, int grid, int block, uint32_t *sink,
                       int *ops_per_thread_iter, slamfe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLAMFE_H */
