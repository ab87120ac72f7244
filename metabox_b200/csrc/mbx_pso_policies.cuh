// mbx_pso_policies.cuh -- arithmetic of the two single-particle PSO agents' rollout policies, written once for the device kernels
// (mbx_pso_policies.cu) and for the host emulation of the CPU tests (tests/emul/pso_policies_host.cpp).
//
// RL_PSO (agent/rl_pso_agent.py:9-45): mu / sigma MLPs 2*dim -> 32 -> 8 -> 1 (ReLU); mu = (tanh + 1) / 2, sigma = clamp((tanh + 1) / 2,
// 0.01, 0.7); a = mu + sigma * eps; a sample with |a - 0.5| >= 0.5 becomes (a + 3 sigma - mu) * (1/6 * sigma) (:33-34, as written there).
// Packed weights per network (mu first, then sigma; agent.py RL_PSO_PolicyNet.pack()):
//   L0: W^T [2*dim in][32 out] | b [32]     L1: W [8 out][32 in] | b [8]     L2: w [8] | b [1]
// QLPSO (agent/qlpso_agent.py:17-37): action ~ softmax(Q[state]) over 4 neighbourhood sizes, sampled by inverse CDF from a uniform u
// (agent.py QLPSO_QTable.forward: the first k with u < cdf[k], at most 3).
#pragma once
#include "mbx_hostdev.cuh"

namespace rpp {
constexpr int H0 = 32, H1 = 8;
constexpr float MAX_SIGMA = 0.7f, MIN_SIGMA = 0.01f;
MBX_HD int net_size(int n_in) { return n_in * H0 + H0 + H1 * H0 + H1 + H1 + 1; }
MBX_HD int n_weights(int n_in) { return 2 * net_size(n_in); }

// first hidden layer, neuron j: relu(b0[j] + sum_i W0[j][i] * x[i])
MBX_HD float hidden0(const float* N, int n_in, const float* x, int j) {
    float a = N[n_in * H0 + j];
    for (int i = 0; i < n_in; ++i) a = fmaf(N[i * H0 + j], x[i], a);
    return fmaxf(a, 0.f);
}
// second hidden layer, neuron k (0..7) from the 32 activations
MBX_HD float hidden1(const float* N, int n_in, const float* h0, int k) {
    const float* W = N + n_in * H0 + H0;
    float a = W[H1 * H0 + k];
#pragma unroll
    for (int i = 0; i < H0; ++i) a = fmaf(W[k * H0 + i], h0[i], a);
    return fmaxf(a, 0.f);
}
MBX_HD float output(const float* N, int n_in, const float* h1) {
    const float* W = N + n_in * H0 + H0 + H1 * H0 + H1;
    float a = W[H1];
#pragma unroll
    for (int k = 0; k < H1; ++k) a = fmaf(W[k], h1[k], a);
    return a;
}
MBX_HD float to_mu(float o) { return (tanhf(o) + 1.f) * 0.5f; }
MBX_HD float to_sigma(float o) { return fminf(fmaxf((tanhf(o) + 1.f) * 0.5f, MIN_SIGMA), MAX_SIGMA); }
MBX_HD float sample_action(float mu, float sigma, float eps) {
    const float a = fmaf(sigma, eps, mu);
    return fabsf(a - 0.5f) >= 0.5f ? (a + 3.f * sigma - mu) * (0.16666667f * sigma) : a;
}
}  // namespace rpp

namespace qlp {
constexpr int NS = 4, NA = 4;
// q [4][4] f64, state in 0..3 (anything else is clamped), u in [0, 1)
MBX_HD int sample(const double* q, int state, double u) {
    state = state < 0 ? 0 : (state >= NS ? NS - 1 : state);
    const double* r = q + state * NA;
    double m = r[0];
#pragma unroll
    for (int k = 1; k < NA; ++k) m = r[k] > m ? r[k] : m;
    double e[NA], tot = 0.;
#pragma unroll
    for (int k = 0; k < NA; ++k) { e[k] = exp(r[k] - m); tot += e[k]; }
    double cdf = 0.;
    int a = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) { cdf += e[k] / tot; a += (u >= cdf) ? 1 : 0; }
    return a > NA - 1 ? NA - 1 : a;
}
}  // namespace qlp
