// Host emulation of rl_pso_policy_kernel / qlpso_policy_kernel (metabox_b200/csrc/mbx_pso_policies.cu) for the CPU test-suite: the
// arithmetic is the SAME code the device runs (mbx_pso_policies.cuh compiled by g++); loops replace threads, arrays the warp shuffles.
// TEST INFRASTRUCTURE ONLY, never linked into libmbx.so.
#include <cstddef>
#include <cstdint>

#include "../../metabox_b200/csrc/mbx_pso_policies.cuh"

extern "C" int rpp_host_weight_count(int n_in) { return rpp::n_weights(n_in); }

extern "C" int rpp_host_forward(const float* state, const float* weights, const float* noise, float* action, float* mu_out, float* sigma_out,
                                int B, int n_in) {
    const float* Nm = weights;
    const float* Ns = weights + rpp::net_size(n_in);
    for (int lane = 0; lane < B; ++lane) {
        const float* x = state + (size_t)lane * n_in;
        float h0m[rpp::H0], h0s[rpp::H0], h1m[rpp::H1], h1s[rpp::H1];
        for (int j = 0; j < rpp::H0; ++j) { h0m[j] = rpp::hidden0(Nm, n_in, x, j); h0s[j] = rpp::hidden0(Ns, n_in, x, j); }
        for (int k = 0; k < rpp::H1; ++k) { h1m[k] = rpp::hidden1(Nm, n_in, h0m, k); h1s[k] = rpp::hidden1(Ns, n_in, h0s, k); }
        const float mu = rpp::to_mu(rpp::output(Nm, n_in, h1m)), sigma = rpp::to_sigma(rpp::output(Ns, n_in, h1s));
        action[lane] = rpp::sample_action(mu, sigma, noise ? noise[lane] : 0.f);
        if (mu_out) mu_out[lane] = mu;
        if (sigma_out) sigma_out[lane] = sigma;
    }
    return 0;
}

extern "C" int qlp_host_forward(const int32_t* q_state, const double* q_table, const float* u, int32_t* action, int B) {
    for (int lane = 0; lane < B; ++lane) action[lane] = qlp::sample(q_table, q_state[lane], (double)u[lane]);
    return 0;
}
