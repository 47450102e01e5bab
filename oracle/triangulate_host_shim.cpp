// TEST INFRASTRUCTURE ONLY.  Host build of the product's link-triangulation arithmetic
// (csrc/triangulate_core.cuh) for the CPU test suite; build with -ffp-contract=off is NOT wanted here:
// nvcc contracts a*b-c into FMAs on the device, so the host build uses -ffp-contract=fast to follow it
// (results are compared with a tolerance, not bitwise).  The product never loads this library.
#include "../67604-slam---video-navigation_b200/csrc/triangulate_core.cuh"

extern "C" void triangulate_host_links(const double *links, long n, const double *P, const double *Q, double *xyz)
{
    slamfe::Cams c;
    for (int k = 0; k < 12; ++k) { c.P[k] = P[k]; c.Q[k] = Q[k]; }
    for (long i = 0; i < n; ++i)
        slamfe::triangulate_link(c, links[3 * i], links[3 * i + 1], links[3 * i + 2], xyz[3 * i], xyz[3 * i + 1],
                                 xyz[3 * i + 2]);
}
