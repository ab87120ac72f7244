// Host emulation of deddqn_policy_kernel (metabox_b200/csrc/mbx_deddqn_policy.cu) for the CPU test-suite: the per-neuron arithmetic is the
// SAME code the device runs (mbx_deddqn_policy.cuh compiled by g++); the tile loop, the staging of the transposed activations and the layer
// order are restated here in the kernel's order.  TEST INFRASTRUCTURE ONLY, never linked into libmbx.so.
#include <cstdint>
#include <vector>

#include "../../metabox_b200/csrc/mbx_deddqn_policy.cuh"

extern "C" int ddq_host_weight_count(void) { return ddq::N_WEIGHTS; }

extern "C" int ddq_host_forward(const float* state, const float* weights, int32_t* action, float* q_out, int B) {
    std::vector<float> A(ddq::NHID * ddq::TL), Bf(ddq::NHID * ddq::TL);
    float Q[ddq::TL * ddq::NACT];
    for (int lane0 = 0; lane0 < B; lane0 += ddq::TL) {
        for (int idx = 0; idx < ddq::TL * ddq::NIN; ++idx) {
            const int l = idx / ddq::NIN, i = idx - l * ddq::NIN, lane = lane0 + l;
            A[i * ddq::TL + l] = lane < B ? ddq::sanitise(state[(size_t)lane * ddq::NIN + i]) : 0.f;
        }
        for (int o = 0; o < ddq::NHID; ++o) ddq::hidden_neuron(weights + ddq::O_L0, ddq::NIN, A.data(), o, Bf.data());
        for (int o = 0; o < ddq::NHID; ++o) ddq::hidden_neuron(weights + ddq::O_L1, ddq::NHID, Bf.data(), o, A.data());
        for (int o = 0; o < ddq::NHID; ++o) ddq::hidden_neuron(weights + ddq::O_L2, ddq::NHID, A.data(), o, Bf.data());
        for (int o = 0; o < ddq::NHID; ++o) ddq::hidden_neuron(weights + ddq::O_L3, ddq::NHID, Bf.data(), o, A.data());
        for (int t = 0; t < ddq::TL * ddq::NACT; ++t) Q[t] = ddq::q_value(weights + ddq::O_OUT, A.data(), t % ddq::NACT, t / ddq::NACT);
        for (int t = 0; t < ddq::TL && lane0 + t < B; ++t) {
            action[lane0 + t] = ddq::greedy(Q + t * ddq::NACT);
            if (q_out)
                for (int a = 0; a < ddq::NACT; ++a) q_out[(size_t)(lane0 + t) * ddq::NACT + a] = Q[t * ddq::NACT + a];
        }
    }
    return 0;
}
