/*
 * ORACLE — test infrastructure only.  NOT part of the product path.
 *
 * Plain-C restatement of the brute-force Hamming matcher semantics the reference obtains
 * from OpenCV (third-party, not vendored under /root/reference; image has opencv-python
 * 4.13.0, the reference pins no version):
 *
 *   cv2.BFMatcher(NORM_HAMMING, crossCheck=False).match(q, t)
 *        call sites: final_project/backend/database/database.py:54-55,
 *                    final_project/backend/loop/loop_closure.py:422,
 *                    final_project/algorithms/matching.py:15
 *   cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(l, r)
 *        call site:  final_project/algorithms/matching.py:44 (factory matching.py:19-24)
 *   knnMatch(q, t, k=2) + ratio test
 *        call sites: VAN_ex/code/ex1.py:189-190, ex1.py:118-122 (GOOD_RATIO=0.6, ex1.py:9)
 *
 * Published algorithm (OpenCV BFMatcher / batchDistance): for every query row compute the
 * Hamming distance (popcount of XOR over the descriptor bytes) to every train row, keep the k
 * smallest in ascending (distance, trainIdx) order — i.e. the FIRST minimum wins ties.
 * crossCheck keeps (i, j) iff j is the first-min of row i AND i is the first-min of column j.
 * Pinned against cv2 4.13.0 outputs by tests/test_oracle.py via tests/golden/*.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static inline int hamming_row(const uint8_t *a, const uint8_t *b, int nbytes)
{
    int d = 0, k = 0;
    for (; k + 8 <= nbytes; k += 8) {
        uint64_t x, y;
        memcpy(&x, a + k, 8);
        memcpy(&y, b + k, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; k < nbytes; ++k)
        d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
    return d;
}

/*
 * Top-2 nearest train rows for every query row, ascending (distance, index).
 * idx2/dist2 are (nq, 2) int32; missing neighbours (nt < 2) are -1.
 */
void oracle_hamming_top2(const uint8_t *q, int nq, int q_stride,
                         const uint8_t *t, int nt, int t_stride,
                         int nbytes, int32_t *idx2, int32_t *dist2)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        int b1 = 1 << 30, b2 = 1 << 30, i1 = -1, i2 = -1;
        const uint8_t *qi = q + (size_t)i * q_stride;
        for (int j = 0; j < nt; ++j) {
            int d = hamming_row(qi, t + (size_t)j * t_stride, nbytes);
            if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
            else if (d < b2) { b2 = d; i2 = j; }
        }
        idx2[2 * i] = i1;      dist2[2 * i] = i1 >= 0 ? b1 : -1;
        idx2[2 * i + 1] = i2;  dist2[2 * i + 1] = i2 >= 0 ? b2 : -1;
    }
}

/*
 * First-min of every COLUMN of the distance matrix (= 1-NN of each train row among the
 * queries, lowest query index on ties).  Used for crossCheck and for the backward match of
 * database.py:55.
 */
void oracle_hamming_colmin(const uint8_t *q, int nq, int q_stride,
                           const uint8_t *t, int nt, int t_stride,
                           int nbytes, int32_t *idx, int32_t *dist)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nt; ++j) {
        int b = 1 << 30, bi = -1;
        const uint8_t *tj = t + (size_t)j * t_stride;
        for (int i = 0; i < nq; ++i) {
            int d = hamming_row(q + (size_t)i * q_stride, tj, nbytes);
            if (d < b) { b = d; bi = i; }
        }
        idx[j] = bi;
        dist[j] = bi >= 0 ? b : -1;
    }
}

/* Full distance matrix (small cases only): D is (nq, nt) int32. */
void oracle_hamming_matrix(const uint8_t *q, int nq, int q_stride,
                           const uint8_t *t, int nt, int t_stride,
                           int nbytes, int32_t *D)
{
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < nt; ++j)
            D[(size_t)i * nt + j] =
                hamming_row(q + (size_t)i * q_stride, t + (size_t)j * t_stride, nbytes);
}

int oracle_version(void) { return 1; }
