// mbx_pso_policies.cu -- rollout policies of the two single-particle PSO agents as ONE kernel each (their backbones advance one particle
// per environment step, so the ~20 / ~9 torch launches of the policy between two fused steps are a large part of a step):
//   rl_pso_policy_kernel : state [B, 2*dim] -> action [B]       (agent/rl_pso_agent.py:22-45, rollout :116-124); one warp per lane
//   qlpso_policy_kernel  : q_state [B] -> action [B] in 0..3     (agent/qlpso_agent.py:33-36, rollout :66-75);   one thread per lane
#include "mbx_common.cuh"
#include "mbx_launch.h"
#include "mbx_pso_policies.cuh"

#define PSP_NT 128
#define PSP_MAX_IN 64

__global__ void __launch_bounds__(PSP_NT)
rl_pso_policy_kernel(const float* __restrict__ state, const float* __restrict__ weights, const float* __restrict__ noise,
                     float* __restrict__ action, float* __restrict__ mu_out, float* __restrict__ sigma_out, int B, int n_in) {
    __shared__ float sx[PSP_NT / 32][PSP_MAX_IN];
    const int w = threadIdx.x >> 5, j = threadIdx.x & 31, lane = blockIdx.x * (PSP_NT / 32) + w;
    if (lane >= B) return;                                           // whole warps leave together; no block-wide barrier below
    for (int i = j; i < n_in; i += 32) sx[w][i] = state[(size_t)lane * n_in + i];
    __syncwarp();
    const float* Nm = weights;
    const float* Ns = weights + rpp::net_size(n_in);
    const float am = rpp::hidden0(Nm, n_in, sx[w], j), as = rpp::hidden0(Ns, n_in, sx[w], j);
    float h0m[rpp::H0], h0s[rpp::H0];
#pragma unroll
    for (int k = 0; k < rpp::H0; ++k) { h0m[k] = __shfl_sync(0xffffffffu, am, k); h0s[k] = __shfl_sync(0xffffffffu, as, k); }
    const float bm = rpp::hidden1(Nm, n_in, h0m, j & 7), bs = rpp::hidden1(Ns, n_in, h0s, j & 7);
    float h1m[rpp::H1], h1s[rpp::H1];
#pragma unroll
    for (int k = 0; k < rpp::H1; ++k) { h1m[k] = __shfl_sync(0xffffffffu, bm, k); h1s[k] = __shfl_sync(0xffffffffu, bs, k); }
    if (j == 0) {
        const float mu = rpp::to_mu(rpp::output(Nm, n_in, h1m)), sigma = rpp::to_sigma(rpp::output(Ns, n_in, h1s));
        action[lane] = rpp::sample_action(mu, sigma, noise ? noise[lane] : 0.f);
        if (mu_out) mu_out[lane] = mu;
        if (sigma_out) sigma_out[lane] = sigma;
    }
}

__global__ void __launch_bounds__(PSP_NT)
qlpso_policy_kernel(const int32_t* __restrict__ q_state, const double* __restrict__ q_table, const float* __restrict__ u,
                    int32_t* __restrict__ action, int B) {
    __shared__ double sq[qlp::NS * qlp::NA];
    if (threadIdx.x < qlp::NS * qlp::NA) sq[threadIdx.x] = q_table[threadIdx.x];
    __syncthreads();
    const int lane = blockIdx.x * PSP_NT + threadIdx.x;
    if (lane < B) action[lane] = qlp::sample(sq, q_state[lane], (double)u[lane]);
}

int mbx_launch_rl_pso_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                             int B, int n_in, cudaStream_t st) {
    constexpr int LPC = PSP_NT / 32;
    rl_pso_policy_kernel<<<(B + LPC - 1) / LPC, PSP_NT, 0, st>>>(state, weights, noise, action, mu, sigma, B, n_in);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

int mbx_launch_qlpso_policy(const int32_t* q_state, const double* q_table, const float* u, int32_t* action, int B, cudaStream_t st) {
    qlpso_policy_kernel<<<(B + PSP_NT - 1) / PSP_NT, PSP_NT, 0, st>>>(q_state, q_table, u, action, B);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
