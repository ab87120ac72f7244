// mbx_walk.cu -- RL_HPSDE's progressive random walk for B lanes: one thread per (lane, dimension) follows its coordinate through the
// n_steps sequential moves (optimizer/rl_hpsde_optimizer.py:150-167); the draws come from the caller (torch generator or a recorded tape).
#include "mbx_common.cuh"
#include "mbx_launch.h"
#include "mbx_walk.cuh"

__global__ void __launch_bounds__(128)
progressive_walk_kernel(const double* __restrict__ zone_u, const double* __restrict__ r_u, const int32_t* __restrict__ rD,
                        const double* __restrict__ step_u, double lb, double ub, double step_size, int B, int D, int n_steps,
                        double* __restrict__ out, float* __restrict__ out32) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * D) return;
    const int b = e / D, d = e - b * D;
    pwalk::coordinate(lb, ub, step_size, zone_u[e], r_u[e], rD[b] == d, step_u + (size_t)b * n_steps * D + d, n_steps, D,
                      out + (size_t)b * (n_steps + 1) * D + d, out32 ? out32 + (size_t)b * (n_steps + 1) * D + d : nullptr);
}

int mbx_launch_progressive_walk(const double* zone_u, const double* r_u, const int32_t* rD, const double* step_u, double lb, double ub,
                                double step_size, int B, int D, int n_steps, double* out, float* out32, cudaStream_t st) {
    const int n = B * D;
    progressive_walk_kernel<<<(n + 127) / 128, 128, 0, st>>>(zone_u, r_u, rD, step_u, lb, ub, step_size, B, D, n_steps, out, out32);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
