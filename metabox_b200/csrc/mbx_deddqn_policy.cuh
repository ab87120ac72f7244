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

// The same two functions for ONE lane with its activations as a plain vector x[i] (the persistent kernel of mbx_deddqn_run.cu: a lane group
// evaluates its own lane).  Same chain as above -- bias first, inputs in ascending order, one fmaf each -- hence the same bits.
MBX_HD float hidden_neuron1(const float* L, int n_in, const float* x, int o) {
    float a = L[n_in * NHID + o];
#pragma unroll 10
    for (int i = 0; i < n_in; ++i) a = fmaf(L[i * NHID + o], x[i], a);
    return fmaxf(a, 0.f);
}
MBX_HD float q_value1(const float* L, const float* x, int a) {
    float q = L[NACT * NHID + a];
#pragma unroll 10
    for (int i = 0; i < NHID; ++i) q = fmaf(L[a * NHID + i], x[i], q);
    return q;
}

#if defined(__CUDACC__)
// Hidden layer for the TL = 8 lanes of a CTA of the persistent kernel (mbx_deddqn_run.cu): activations transposed xT[i][8], packed
// fma.rn.f32x2 -- each component is the IEEE fma of `hidden_neuron`'s scalar chain (bias first, inputs in ascending order), so the bits
// are the same.  The loop is bound by load LATENCY (weights from L2 / L1, ~300 cycles; activations from shared memory, ~30), not by
// bandwidth: what matters is how many loads are in flight.
// Work item = (neuron o, half h of the 8 lanes): 200 threads per layer, per input one weight load + one LDS.128 (lanes 4h .. 4h+3) + two
// packed fmas.  The weights of the NEXT chunk of CH inputs are in flight (registers) while the current chunk is consumed, activations are
// read BLK inputs ahead.
template <int N_IN, int CH, int BLK>
__device__ __noinline__ void hidden_neuron_x8(const float* __restrict__ L, const float* xT, const int o, const int h, float* yT) {
    static_assert(CH % BLK == 0, "whole activation blocks per weight chunk");
    constexpr int NB = (N_IN + BLK - 1) / BLK;                            // activation blocks
    const float* Lo = L + o;
    const float bias = __ldg(Lo + N_IN * NHID);
    float2 a01 = make_float2(bias, bias), a23 = a01;
    const float* xh = xT + 4 * h;
    float wc[CH], wn[CH];
    float4 xa[BLK], xb[BLK];
    // every index below is a compile-time constant (all loops unrolled): no address arithmetic, no tail clamps.  Software pipeline: the
    // weights of chunk c + 1 and the activations of block b + 1 are requested before the fmas of block b are issued.
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (k < N_IN) wc[k] = __ldg(Lo + k * NHID);
#pragma unroll
    for (int k = 0; k < BLK; ++k) xa[k] = *reinterpret_cast<const float4*>(xh + k * TL);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i0 = b * BLK;
        if (i0 % CH == 0) {                                                // first block of a chunk: request the next chunk's weights
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (i0 + CH + k < N_IN) wn[k] = __ldg(Lo + (i0 + CH + k) * NHID);
        }
        float4* xc = (b & 1) ? xb : xa;
        float4* xn = (b & 1) ? xa : xb;
#pragma unroll
        for (int k = 0; k < BLK; ++k)
            if (i0 + BLK + k < N_IN) xn[k] = *reinterpret_cast<const float4*>(xh + (i0 + BLK + k) * TL);
#pragma unroll
        for (int k = 0; k < BLK; ++k) {
            if (i0 + k < N_IN) {
                const float wv = wc[(i0 + k) % CH];
                const float2 ww = make_float2(wv, wv);
                a01 = __ffma2_rn(ww, make_float2(xc[k].x, xc[k].y), a01);
                a23 = __ffma2_rn(ww, make_float2(xc[k].z, xc[k].w), a23);
            }
        }
        if ((i0 + BLK) % CH == 0) {                                        // last block of a chunk
#pragma unroll
            for (int k = 0; k < CH; ++k) wc[k] = wn[k];
        }
    }
    *reinterpret_cast<float4*>(yT + o * TL + 4 * h) = make_float4(fmaxf(a01.x, 0.f), fmaxf(a01.y, 0.f), fmaxf(a23.x, 0.f), fmaxf(a23.y, 0.f));
}
// Q(lane l, action a) on the transposed activations: the chain of `q_value`, its loads issued a quarter of the inputs at a time
__device__ __forceinline__ float q_value_x8(const float* __restrict__ L, const float* xT, const int a, const int l) {
    float q = __ldg(L + NACT * NHID + a);
#pragma unroll 1
    for (int i0 = 0; i0 < NHID; i0 += NHID / 4) {
        float w[NHID / 4], x[NHID / 4];
#pragma unroll
        for (int k = 0; k < NHID / 4; ++k) { w[k] = __ldg(L + a * NHID + i0 + k); x[k] = xT[(i0 + k) * TL + l]; }
#pragma unroll
        for (int k = 0; k < NHID / 4; ++k) q = fmaf(w[k], x[k], q);
    }
    return q;
}
#endif

// torch.argmax over the 4 actions: the first maximum
MBX_HD int greedy(const float* q) {
    int best = 0;
#pragma unroll
    for (int a = 1; a < NACT; ++a) best = q[a] > q[best] ? a : best;
    return best;
}
}  // namespace ddq
