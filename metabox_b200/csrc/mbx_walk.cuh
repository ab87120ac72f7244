// mbx_walk.cuh -- one coordinate of RL_HPSDE's progressive random walk (optimizer/rl_hpsde_optimizer.py:150-167), written once for the
// device kernel (mbx_walk.cu: one thread per (lane, dimension)) and for the host emulation of the CPU tests (tests/emul/walk_host.cpp).
// The coordinates of a walk are independent of each other: the start lies in the half of the box chosen by `zone_u` (on the box face for
// the one dimension `rD`), every step moves by U(0,1) * step_size against the current zone sign, a point that leaves the box is mirrored
// ONCE at the face it crossed (it may stay outside when the box is narrower than the step: the protein box) and the zone sign flips.
#pragma once
#include "mbx_hostdev.cuh"

namespace pwalk {
// step_u / out are strided by `stride` elements per walk step; out has n_steps + 1 points
MBX_HD void coordinate(double lb, double ub, double step_size, double zone_u, double r_u, bool on_face, const double* step_u, int n_steps,
                       long stride, double* out, float* out32) {
    double zone = zone_u < 0.5 ? -1.0 : 1.0;
    double s = (ub + lb) / 2 + zone * (r_u * (ub - lb) / 2);
    if (on_face) s = zone == -1.0 ? lb : ub;
    out[0] = s;
    if (out32) out32[0] = (float)s;
    for (int k = 1; k <= n_steps; ++k) {
        s = s + step_u[(k - 1) * stride] * (-step_size) * zone;
        const bool hi = s > ub, lo = s < lb;
        if (hi) s = 2 * ub - s;
        if (lo) s = 2 * lb - s;
        if (hi || lo) zone = -zone;
        out[k * stride] = s;
        if (out32) out32[k * stride] = (float)s;
    }
}
}  // namespace pwalk
