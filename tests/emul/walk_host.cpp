// Host emulation of progressive_walk_kernel (metabox_b200/csrc/mbx_walk.cu): the same per-coordinate function (mbx_walk.cuh compiled by g++),
// a loop over (lane, dimension) instead of threads.  TEST INFRASTRUCTURE ONLY, never linked into libmbx.so.
#include <cstddef>
#include <cstdint>

#include "../../metabox_b200/csrc/mbx_walk.cuh"

extern "C" int pwalk_host(const double* zone_u, const double* r_u, const int32_t* rD, const double* step_u, double lb, double ub,
                          double step_size, int B, int D, int n_steps, double* out, float* out32) {
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            pwalk::coordinate(lb, ub, step_size, zone_u[b * D + d], r_u[b * D + d], rD[b] == d, step_u + (size_t)b * n_steps * D + d, n_steps, D,
                              out + (size_t)b * (n_steps + 1) * D + d, out32 ? out32 + (size_t)b * (n_steps + 1) * D + d : nullptr);
    return 0;
}
