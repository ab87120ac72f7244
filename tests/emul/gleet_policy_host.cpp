// Host emulation of gleet_policy_kernel (metabox_b200/csrc/mbx_gleet_policy.cu) for the CPU test-suite: the per-particle arithmetic is
// the SAME code the device runs (mbx_gleet_policy.cuh compiled by g++); only the orchestration -- one loop iteration per thread, a plain
// sum where the kernel does a block reduction, nothing where it has a barrier -- is restated here, phase by phase in the kernel's order.
// TEST INFRASTRUCTURE ONLY: built by tests/test_host_cpu.py into a temporary directory, never linked into libmbx.so.
#include <vector>

#include "../../metabox_b200/csrc/mbx_gleet_policy.cuh"

namespace {
struct Particle { float h[glp::E], din[glp::E], qs[glp::E], x[glp::E]; };

void swarm_norm(std::vector<Particle>& P, int NP) {
    float s = 0.f;
    for (int p = 0; p < NP; ++p) s += glp::sum16(P[p].x);
    const float mean = s / (float)(NP * glp::E);
    float q = 0.f;
    for (int p = 0; p < NP; ++p) q += glp::sqdev16(P[p].x, mean);
    const float inv = glp::inv_std(q, NP);
    for (int p = 0; p < NP; ++p) glp::normalise(P[p].x, mean, inv);
}

void block(const float* Wb, bool decoder, std::vector<Particle>& P, int NP, float* K, float* V) {
    for (int p = 0; p < NP; ++p) glp::project(Wb, decoder ? P[p].din : P[p].h, P[p].h, P[p].qs, K + p * glp::E, V + p * glp::E);
    for (int p = 0; p < NP; ++p) glp::attend(Wb, K, V, NP, P[p].qs, P[p].h, P[p].x);
    swarm_norm(P, NP);
    for (int p = 0; p < NP; ++p) glp::feed_forward(Wb, P[p].x);
    swarm_norm(P, NP);
    for (int p = 0; p < NP; ++p)
        for (int c = 0; c < glp::E; ++c) P[p].h[c] = P[p].x[c];
}
}  // namespace

extern "C" int glp_host_weight_count(void) { return glp::N_WEIGHTS; }

extern "C" int glp_host_forward(const float* state, const float* weights, const float* noise, float* action, float* mu_out, float* sigma_out,
                                float* logits_out, int B, int NP) {
    if (NP < 1 || NP > 128) return -1;
    std::vector<Particle> P(NP);
    std::vector<float> K((size_t)NP * glp::E), V((size_t)NP * glp::E);
    for (int b = 0; b < B; ++b) {
        for (int p = 0; p < NP; ++p) glp::embed(weights, state + ((size_t)b * NP + p) * glp::NF, P[p].h, P[p].din);
        block(weights + glp::O_ENC, false, P, NP, K.data(), V.data());
        block(weights + glp::O_DEC, true, P, NP, K.data(), V.data());
        for (int p = 0; p < NP; ++p) {
            const size_t o = (size_t)b * NP + p;
            float mu, sigma;
            glp::mu_sigma(weights, P[p].h, mu, sigma);
            action[o] = glp::sample_action(mu, sigma, noise ? noise[o] : 0.f);
            if (mu_out) mu_out[o] = mu;
            if (sigma_out) sigma_out[o] = sigma;
            if (logits_out)
                for (int c = 0; c < glp::E; ++c) logits_out[o * glp::E + c] = P[p].h[c];
        }
    }
    return 0;
}
