// mbx_rlepso_policy.cu -- the RLEPSO actor's rollout forward as ONE kernel: state [B, 1] -> action [B, 35].
// Reference: agent/rlepso_agent.py:22-47 (Actor.forward), :294-303 (rollout loop: one action per generation).
//
// The torch module costs ~20 small launches per generation (two 3-layer MLPs, tanh / scale / sample / clamp); in the striped sweep that is
// the 4.7 % between the step-kernel-only rate and the measured headline.  Here one warp owns one lane: thread j
// computes hidden neuron j of both networks (the first layer, 64 units of one input, is recomputed on the fly), the 32 activations
// travel by warp shuffles, thread j then produces outputs j and 32 + j.  Weights (26.9 KB) are read through L1 / L2 by every warp.
#include "mbx_common.cuh"
#include "mbx_launch.h"
#include "mbx_rlepso_policy.cuh"

#define RLP_NT 128

__global__ void __launch_bounds__(RLP_NT)
rlepso_policy_kernel(const float* __restrict__ state, const float* __restrict__ weights, const float* __restrict__ noise,
                     float* __restrict__ action, float* __restrict__ mu_out, float* __restrict__ sigma_out, int B) {
    const int lane = blockIdx.x * (RLP_NT / 32) + (threadIdx.x >> 5), j = threadIdx.x & 31;
    if (lane >= B) return;                                           // whole warps leave together
    const float x = state[lane];
    const float* Nm = weights;
    const float* Ns = weights + rlp::NET_SIZE;
    const float hm = rlp::hidden1(Nm, x, j), hs = rlp::hidden1(Ns, x, j);
    float h1m[rlp::H1], h1s[rlp::H1];
#pragma unroll
    for (int k = 0; k < rlp::H1; ++k) {
        h1m[k] = __shfl_sync(0xffffffffu, hm, k);
        h1s[k] = __shfl_sync(0xffffffffu, hs, k);
    }
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int o = rep * 32 + j;
        if (o < rlp::NA) {
            const float mu = rlp::to_mu(rlp::output(Nm, h1m, o)), sigma = rlp::to_sigma(rlp::output(Ns, h1s, o));
            const size_t at = (size_t)lane * rlp::NA + o;
            action[at] = rlp::sample_action(mu, sigma, noise ? noise[at] : 0.f);
            if (mu_out) mu_out[at] = mu;
            if (sigma_out) sigma_out[at] = sigma;
        }
    }
}

int mbx_launch_rlepso_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                             int B, cudaStream_t st) {
    constexpr int LPC = RLP_NT / 32;
    rlepso_policy_kernel<<<(B + LPC - 1) / LPC, RLP_NT, 0, st>>>(state, weights, noise, action, mu, sigma, B);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
