// Shared device code of the link triangulation (used by triangulate.cu and tracking.cu).
#pragma once
#include "common.cuh"
#include "triangulate_core.cuh"

namespace slamfe {

inline int load_cams(const double *P, const double *Q, Cams &c)
{
    if (!P || !Q) return SLAMFE_EINVAL;
    for (int k = 0; k < 12; ++k) {
        c.P[k] = P[k];
        c.Q[k] = Q[k];
    }
    return 0;
}

}  // namespace slamfe
