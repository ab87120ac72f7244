// mbx_deddqn_policy.cu -- the DE-DDQN agent's greedy rollout action as ONE kernel: state [B, 99] -> action [B] (argmax of the Q-network).
// Reference: agent/de_ddqn_agent.py:21-25 (network), :59-68 (`__get_action`), :108-117 (rollout loop: one action per function evaluation).
//
// Why a kernel: DE-DDQN spends one environment step per function evaluation (59 900 serial steps per run); between two fused step kernels
// the torch Q-network costs ~12 launches (nan_to_num, 5 addmm, 4 relu, argmax, cast), about as long as the step itself in the unstriped
// sweep (profiles/r1h_bench_n1_deddqn.json: kernel_share_of_step 0.51).  Here a CTA owns a tile of 8 lanes: activations stay in shared
// memory (transposed, two 3.2 KB buffers), thread o computes neuron o for the 8 lanes with one coalesced weight load per input (the 163 KB of
// weights are L2-resident and shared by all CTAs), and the 99-float state rows are read once.
#include "mbx_common.cuh"
#include "mbx_deddqn_policy.cuh"
#include "mbx_launch.h"

#define DDQ_NT 128

__global__ void __launch_bounds__(DDQ_NT)
deddqn_policy_kernel(const float* __restrict__ state, const float* __restrict__ weights, int32_t* __restrict__ action,
                     float* __restrict__ q_out, int B) {
    __shared__ __align__(16) float sA[ddq::NHID * ddq::TL];
    __shared__ __align__(16) float sB[ddq::NHID * ddq::TL];
    __shared__ float sQ[ddq::TL * ddq::NACT];
    const int t = threadIdx.x, lane0 = blockIdx.x * ddq::TL;
    for (int idx = t; idx < ddq::TL * ddq::NIN; idx += DDQ_NT) {
        const int l = idx / ddq::NIN, i = idx - l * ddq::NIN;
        const int lane = lane0 + l;
        sA[i * ddq::TL + l] = lane < B ? ddq::sanitise(state[(size_t)lane * ddq::NIN + i]) : 0.f;
    }
    __syncthreads();
    if (t < ddq::NHID) ddq::hidden_neuron(weights + ddq::O_L0, ddq::NIN, sA, t, sB);
    __syncthreads();
    if (t < ddq::NHID) ddq::hidden_neuron(weights + ddq::O_L1, ddq::NHID, sB, t, sA);
    __syncthreads();
    if (t < ddq::NHID) ddq::hidden_neuron(weights + ddq::O_L2, ddq::NHID, sA, t, sB);
    __syncthreads();
    if (t < ddq::NHID) ddq::hidden_neuron(weights + ddq::O_L3, ddq::NHID, sB, t, sA);
    __syncthreads();
    if (t < ddq::TL * ddq::NACT) sQ[t] = ddq::q_value(weights + ddq::O_OUT, sA, t % ddq::NACT, t / ddq::NACT);      // sQ[l][a]
    __syncthreads();
    if (t < ddq::TL && lane0 + t < B) {
        action[lane0 + t] = ddq::greedy(sQ + t * ddq::NACT);
        if (q_out) {
#pragma unroll
            for (int a = 0; a < ddq::NACT; ++a) q_out[(size_t)(lane0 + t) * ddq::NACT + a] = sQ[t * ddq::NACT + a];
        }
    }
}

int mbx_launch_deddqn_policy(const float* state, const float* weights, int32_t* action, float* q, int B, cudaStream_t st) {
    deddqn_policy_kernel<<<(B + ddq::TL - 1) / ddq::TL, DDQ_NT, 0, st>>>(state, weights, action, q, B);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
