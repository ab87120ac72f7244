// mbx_deddqn_policy.cuh -- per-neuron arithmetic of the DE-DDQN Q-network (agent/de_ddqn_agent.py:21-25: MLP 99 -> 100 -> 100 -> 100 ->
// 100 -> 4, ReLU) and its greedy action (`__get_action`, :59-68, epsilon_greedy=False), written once for the device kernel
// (mbx_deddqn_policy.cu) and for the host emulation of the CPU tests (tests/emul/deddqn_policy_host.cpp).
//
// A CTA advances a TILE of TL lanes; activations live in shared memory TRANSPOSED, xT[i][lane], so that one neuron reads the TL lanes
// of an input as float4s while its weight is a single coalesced global load.  Packed weights (float32, N_WEIGHTS values; agent.py
// DE_DDQN_Net.pack() writes this order):
//   hidden layer k (k = 0..3): W^T [n_in][100] (in-major: neighbouring neurons = neighbouring addresses) | bias [100]
//   output layer:              W   [4][100]    (torch layout)                                            | bias [4]
#pragma once
#include "mbx_hostdev.cuh"

namespace ddq {
using mbxhd::F4;
using mbxhd::ld4;
using mbxhd::st4;
constexpr int NIN = 99, NHID = 100, NACT = 4, TL = 8;
constexpr int O_L0 = 0, O_L1 = O_L0 + NIN * NHID + NHID, O_L2 = O_L1 + NHID * NHID + NHID, O_L3 = O_L2 + NHID * NHID + NHID,
              O_OUT = O_L3 + NHID * NHID + NHID, N_WEIGHTS = O_OUT + NACT * NHID + NACT;
static_assert(O_L1 % 4 == 0 && O_L2 % 4 == 0 && O_L3 % 4 == 0 && O_OUT % 4 == 0 && N_WEIGHTS % 4 == 0, "float4 alignment of the pieces");

// the sweep's torch.nan_to_num(state, nan=0, posinf=0, neginf=0) (features of a finished / degenerate lane)
MBX_HD float sanitise(float v) { return (v == v && fabsf(v) <= 3.402823466e+38f) ? v : 0.f; }

// hidden neuron o of a layer with n_in inputs, for the TL lanes of the tile: yT[o][0..TL) = relu(b[o] + sum_i Wt[i][o] * xT[i][.])
MBX_HD void hidden_neuron(const float* L, int n_in, const float* xT, int o, float* yT) {
    const float* Wt = L;
    const float bias = L[n_in * NHID + o];
    float a[TL];
#pragma unroll
    for (int l = 0; l < TL; ++l) a[l] = bias;
#pragma unroll 4
    for (int i = 0; i < n_in; ++i) {
        const float w = Wt[i * NHID + o];
        const F4 x0 = ld4(xT + i * TL), x1 = ld4(xT + i * TL + 4);
        a[0] = fmaf(w, x0.x, a[0]); a[1] = fmaf(w, x0.y, a[1]); a[2] = fmaf(w, x0.z, a[2]); a[3] = fmaf(w, x0.w, a[3]);
        a[4] = fmaf(w, x1.x, a[4]); a[5] = fmaf(w, x1.y, a[5]); a[6] = fmaf(w, x1.z, a[6]); a[7] = fmaf(w, x1.w, a[7]);
    }
    st4(yT + o * TL, fmaxf(a[0], 0.f), fmaxf(a[1], 0.f), fmaxf(a[2], 0.f), fmaxf(a[3], 0.f));
    st4(yT + o * TL + 4, fmaxf(a[4], 0.f), fmaxf(a[5], 0.f), fmaxf(a[6], 0.f), fmaxf(a[7], 0.f));
}

// Q(lane l, action a) = b[a] + sum_i W[a][i] * xT[i][l]
MBX_HD float q_value(const float* L, const float* xT, int a, int l) {
    float q = L[NACT * NHID + a];
#pragma unroll 4
    for (int i = 0; i < NHID; ++i) q = fmaf(L[a * NHID + i], xT[i * TL + l], q);
    return q;
}

// torch.argmax over the 4 actions: the first maximum
MBX_HD int greedy(const float* q) {
    int best = 0;
#pragma unroll
    for (int a = 1; a < NACT; ++a) best = q[a] > q[best] ? a : best;
    return best;
}
}  // namespace ddq
