// mbx_rlepso_policy.cuh -- per-neuron arithmetic of the RLEPSO actor (agent/rlepso_agent.py:10-47: two MLPs 1 -> 64 -> 32 -> 35 with ReLU,
// mu = (tanh + 1) / 2, sigma = (tanh + 1) / 2 * (0.7 - 0.01) + 0.01, action = clamp(Normal(mu, sigma).sample(), 0, 1)), written once for
// the device kernel (mbx_rlepso_policy.cu: one warp = one lane) and for the host emulation of the CPU tests
// (tests/emul/rlepso_policy_host.cpp).
//
// Packed weights (float32, N_WEIGHTS values; agent.py RLEPSO_Actor.pack() writes this order), per network (mu first, then sigma):
//   L0: w [64] | b [64]            (the state is ONE number: fes / maxFEs)
//   L1: W^T [64 in][32 out] | b [32]
//   L2: W^T [32 in][35 out] | b [35]
#pragma once
#include "mbx_hostdev.cuh"

namespace rlp {
constexpr int H0 = 64, H1 = 32, NA = 35;
constexpr int O_L0W = 0, O_L0B = H0, O_L1W = 2 * H0, O_L1B = O_L1W + H0 * H1, O_L2W = O_L1B + H1, O_L2B = O_L2W + H1 * NA,
              NET_SIZE = O_L2B + NA, N_WEIGHTS = 2 * NET_SIZE;
constexpr float MAX_SIGMA = 0.7f, MIN_SIGMA = 0.01f;

// second hidden layer, neuron j (0..31): relu(b1[j] + sum_i W1[j][i] * relu(w0[i] * x + b0[i]))
MBX_HD float hidden1(const float* N, float x, int j) {
    float a = N[O_L1B + j];
#pragma unroll 8
    for (int i = 0; i < H0; ++i) a = fmaf(N[O_L1W + i * H1 + j], fmaxf(fmaf(N[O_L0W + i], x, N[O_L0B + i]), 0.f), a);
    return fmaxf(a, 0.f);
}
// output o (0..34) from the 32 hidden activations
MBX_HD float output(const float* N, const float* h1, int o) {
    float a = N[O_L2B + o];
#pragma unroll
    for (int k = 0; k < H1; ++k) a = fmaf(N[O_L2W + k * NA + o], h1[k], a);
    return a;
}
MBX_HD float to_mu(float o) { return (tanhf(o) + 1.f) * 0.5f; }
MBX_HD float to_sigma(float o) { return (tanhf(o) + 1.f) * 0.5f * (MAX_SIGMA - MIN_SIGMA) + MIN_SIGMA; }
MBX_HD float sample_action(float mu, float sigma, float eps) { return fminf(fmaxf(fmaf(sigma, eps, mu), 0.f), 1.f); }
}  // namespace rlp
