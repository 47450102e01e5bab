// TEST INFRASTRUCTURE ONLY.  Compiles the product's PnP-refit arithmetic (csrc/refit_core.cuh) for the
// HOST so that the CPU test suite (-m "not gpu") can check the Levenberg-Marquardt refit against an
// independent numpy restatement, against ground truth and against cv2.solvePnP (ransac.py:185-193).
// The product never loads this library.
#include "../67604-slam---video-navigation_b200/csrc/refit_core.cuh"

using namespace slamfe;

// Same control flow as pnp_refit_kernel, one problem, serial sums.  Returns the kernel's status code.
extern "C" int refit_host(const double *T_seed, const double *K, const double *pts, const double *pix,
                          const unsigned char *mask, long n, int max_iter, double tol, double *T_out, double *rms)
{
    double T[12];
    for (int k = 0; k < 12; ++k) T[k] = T_seed[k];
    RefitState st;
    st.lambda = 1e-4;
    st.have_acc = 0;
    long cnt = 0;
    for (long i = 0; i < n; ++i) cnt += mask[i] ? 1 : 0;
    if (cnt < 4) return 0;
    int it = 0, flag = 0;
    for (; it < max_iter; ++it) {
        double acc[REFIT_NACC];
        for (int k = 0; k < REFIT_NACC; ++k) acc[k] = 0.0;
        for (long i = 0; i < n; ++i)
            if (mask[i]) refit_accumulate(T, K, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], pix[2 * i], pix[2 * i + 1], acc);
        double T_try[12];
        flag = refit_step(st, T, acc, T_try, tol);
        if (flag == 0)
            for (int k = 0; k < 12; ++k) T[k] = T_try[k];
        else
            break;
    }
    for (int k = 0; k < 12; ++k) T_out[k] = st.T_acc[k];
    if (rms) *rms = sqrt(st.acc_acc[27] / (double)cnt);
    return flag < 0 ? -1 : (flag == 1 ? it + 1 : -(max_iter + 1));
}
