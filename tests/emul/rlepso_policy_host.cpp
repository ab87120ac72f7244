// Host emulation of rlepso_policy_kernel (metabox_b200/csrc/mbx_rlepso_policy.cu) for the CPU test-suite: the per-neuron arithmetic is the
// SAME code the device runs (mbx_rlepso_policy.cuh compiled by g++); a loop over j replaces the warp's threads and an array the shuffles.
// TEST INFRASTRUCTURE ONLY, never linked into libmbx.so.
#include <cstddef>

#include "../../metabox_b200/csrc/mbx_rlepso_policy.cuh"

extern "C" int rlp_host_weight_count(void) { return rlp::N_WEIGHTS; }

extern "C" int rlp_host_forward(const float* state, const float* weights, const float* noise, float* action, float* mu_out, float* sigma_out,
                                int B) {
    const float* Nm = weights;
    const float* Ns = weights + rlp::NET_SIZE;
    for (int lane = 0; lane < B; ++lane) {
        const float x = state[lane];
        float h1m[rlp::H1], h1s[rlp::H1];
        for (int j = 0; j < rlp::H1; ++j) { h1m[j] = rlp::hidden1(Nm, x, j); h1s[j] = rlp::hidden1(Ns, x, j); }
        for (int o = 0; o < rlp::NA; ++o) {
            const float mu = rlp::to_mu(rlp::output(Nm, h1m, o)), sigma = rlp::to_sigma(rlp::output(Ns, h1s, o));
            const size_t at = (size_t)lane * rlp::NA + o;
            action[at] = rlp::sample_action(mu, sigma, noise ? noise[at] : 0.f);
            if (mu_out) mu_out[at] = mu;
            if (sigma_out) sigma_out[at] = sigma;
        }
    }
    return 0;
}
