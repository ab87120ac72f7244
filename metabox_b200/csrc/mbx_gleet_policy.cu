// mbx_gleet_policy.cu -- the GLEET attention actor's rollout forward as ONE kernel: state [B, NP, 27] -> action [B, NP].
// Reference: agent/gleet_agent.py:386-440 (Actor.forward), :293-303 (rollout loop), agent/networks.py:47-356 (embedder, multi-head
// attention, 'layer' normalisation over the whole swarm, feed-forward sublayer).
//
// Why a kernel: the torch module (agent.GLEET_Actor) launches ~45 small kernels per generation and materialises the
// [B, 4, NP, NP] attention matrices in HBM; measured on B200 it takes 89 % of a GLEET generation
// (profiles/r1i_bench_n1_gleet_actor.json: kernel_share_of_step 0.108).  Here one CTA owns one lane (like the step kernels), one
// thread one particle; the lane's K / V tables (2 x NP x 16 f32) and the 21.7 KB of weights live in shared memory, every
// intermediate in registers, and HBM sees the state once (108 B per particle) and the action once (4 B).
#include "mbx_common.cuh"
#include "mbx_gleet_policy.cuh"
#include "mbx_launch.h"

#define GLP_NT 128

__device__ __forceinline__ float glp_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                   // the previous reduction's readers are done with red[]
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = red[0];
#pragma unroll
    for (int k = 1; k < GLP_NT / 32; ++k) s += red[k];
    return s;
}

// swarm norm of the per-thread 16-vector x (threads p >= NP contribute nothing); two passes like torch.var
__device__ __forceinline__ void glp_swarm_norm(float* x, bool active, int NP, float* red) {
    const float mean = glp_block_sum(active ? glp::sum16(x) : 0.f, red) / (float)(NP * glp::E);
    const float inv = glp::inv_std(glp_block_sum(active ? glp::sqdev16(x, mean) : 0.f, red), NP);
    glp::normalise(x, mean, inv);
}

// one MultiHeadEncoder: queries from qsrc, keys / values and the residual from h; result replaces h
__device__ __forceinline__ void glp_block(const float* Wb, const float* qsrc, float* h, bool active, int p, int NP, float* sK, float* sV,
                                          float* red) {
    float qs[glp::E], x[glp::E];
    __syncthreads();                                   // everybody is done with the previous contents of sK / sV
    if (active) glp::project(Wb, qsrc, h, qs, sK + p * glp::E, sV + p * glp::E);
    __syncthreads();
    if (active) glp::attend(Wb, sK, sV, NP, qs, h, x);
    glp_swarm_norm(x, active, NP, red);
    if (active) glp::feed_forward(Wb, x);
    glp_swarm_norm(x, active, NP, red);
#pragma unroll
    for (int c = 0; c < glp::E; ++c) h[c] = x[c];
}

__global__ void __launch_bounds__(GLP_NT, 4)
gleet_policy_kernel(const float* __restrict__ state, const float* __restrict__ weights, const float* __restrict__ noise,
                    float* __restrict__ action, float* __restrict__ mu_out, float* __restrict__ sigma_out,
                    float* __restrict__ logits_out, int NP) {
    extern __shared__ __align__(16) unsigned char glp_smem[];
    float* sW = reinterpret_cast<float*>(glp_smem);
    float* sK = sW + glp::N_WEIGHTS;
    float* sV = sK + NP * glp::E;
    float* red = sV + NP * glp::E;
    float* sF = sK;                                    // staged state [NP][27] (dead before the first K / V rows are written)
    const int b = blockIdx.x, p = threadIdx.x;
    const bool active = p < NP;
    {
        const float4* w4 = reinterpret_cast<const float4*>(weights);
        float4* s4 = reinterpret_cast<float4*>(sW);
        for (int i = p; i < glp::N_WEIGHTS / 4; i += GLP_NT) s4[i] = w4[i];
        const float* src = state + (size_t)b * NP * glp::NF;
        for (int i = p; i < NP * glp::NF; i += GLP_NT) sF[i] = src[i];
    }
    __syncthreads();
    float h[glp::E], din[glp::E];
    if (active) {
        float f[glp::NF];
#pragma unroll
        for (int i = 0; i < glp::NF; ++i) f[i] = sF[p * glp::NF + i];         // stride 27 words: conflict-free
        glp::embed(sW, f, h, din);
    }
    glp_block(sW + glp::O_ENC, h, h, active, p, NP, sK, sV, red);              // encoder: self-attention over the swarm
    glp_block(sW + glp::O_DEC, din, h, active, p, NP, sK, sV, red);            // decoder: the two 9-feature groups query it
    if (!active) return;
    float mu, sigma;
    glp::mu_sigma(sW, h, mu, sigma);
    const size_t o = (size_t)b * NP + p;
    action[o] = glp::sample_action(mu, sigma, noise ? noise[o] : 0.f);
    if (mu_out) mu_out[o] = mu;
    if (sigma_out) sigma_out[o] = sigma;
    if (logits_out) {
#pragma unroll
        for (int c = 0; c < glp::E; c += 4) glp::st4(logits_out + o * glp::E + c, h[c], h[c + 1], h[c + 2], h[c + 3]);
    }
}

size_t mbx_gleet_policy_smem(int NP) { return sizeof(float) * (glp::N_WEIGHTS + 2 * (size_t)NP * glp::E + 8); }

int mbx_launch_gleet_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                            float* logits, int B, int NP, cudaStream_t st) {
    const size_t smem = mbx_gleet_policy_smem(NP);     // 34.6 KB at NP = 100, 38.1 KB at the maximum NP = 128: below the 48 KB default
    gleet_policy_kernel<<<B, GLP_NT, smem, st>>>(state, weights, noise, action, mu, sigma, logits, NP);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
