// mbx_gleet_policy.cuh -- per-particle arithmetic of the GLEET attention actor (agent/gleet_agent.py:386-440 on
// agent/networks.py:47-356), written once for the device kernel (mbx_gleet_policy.cu: one thread = one particle, one CTA =
// one lane) and for the host emulation the CPU tests compile with g++ (tests/emul/gleet_policy_host.cpp): the same
// functions run in both, only the orchestration (barriers, block reductions) differs.
//
// Packed weights (float32, N_WEIGHTS values; metabox_b200/agent.py GLEET_Actor.pack() writes exactly this order).  Matrices
// whose OUTPUT index is contiguous ("in-major") are read as float4 over outputs, torch Linear matrices ([out][in]) as float4 over
// inputs; every offset is a multiple of 4 floats:
//   O_EMB    [9][16]   embedder.weight^T                      (shared by the three 9-feature groups)
//   O_ENC    block:  WQ | WK | WV  [16 in][16 = head*4 + k]   (W_query[h][e][k] -> [e][h*4+k])
//                    WO  [16 = head*4 + k][16 out]            (W_out.view(-1, 16))
//                    F1W [16 out][16 in] | F1B [16] | F2W [16][16] | F2B [16]
//   O_DEMB   [16 out][32 in]   embedder_for_decoder.weight
//   O_DEC    block like O_ENC
//   O_MU / O_SIGMA   head: L0W [32][16] | L0B [32] | L1W [8][32] | L1B [8] | L2W [8] | L2B [1] (+3 pad)
#pragma once
#include "mbx_hostdev.cuh"
#define GLP_HD MBX_HD

namespace glp {
using mbxhd::F4;
using mbxhd::ld4;
using mbxhd::st4;
constexpr int E = 16, ND = 9, NF = 27, NH = 4, HK = 4;
constexpr int BLK_WQ = 0, BLK_WK = 256, BLK_WV = 512, BLK_WO = 768, BLK_F1W = 1024, BLK_F1B = 1280, BLK_F2W = 1296, BLK_F2B = 1552,
              BLK_SIZE = 1568;
constexpr int HD_L0W = 0, HD_L0B = 512, HD_L1W = 544, HD_L1B = 800, HD_L2W = 808, HD_L2B = 816, HD_SIZE = 820;
constexpr int O_EMB = 0, O_ENC = 144, O_DEMB = O_ENC + BLK_SIZE, O_DEC = O_DEMB + 512, O_MU = O_DEC + BLK_SIZE,
              O_SIGMA = O_MU + HD_SIZE, N_WEIGHTS = O_SIGMA + HD_SIZE;
constexpr float MAX_SIGMA = 0.7f, MIN_SIGMA = 0.01f, LEAKY = 0.01f, NORM_EPS = 1e-05f;
constexpr float QSCALE = 0.5f * 1.4426950408889634f;          // 1/sqrt(key_dim) * log2(e): softmax through exp2f

// y[o] (+)= sum_i x[i] * W[i][o]      W in-major [NI][NO]
template <int NI, int NO, bool ACC>
GLP_HD void mv_in(const float* W, const float* x, float* y) {
    if (!ACC) {
#pragma unroll
        for (int o = 0; o < NO; ++o) y[o] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
        for (int o4 = 0; o4 < NO / 4; ++o4) {
            const F4 w = ld4(W + i * NO + 4 * o4);
            y[4 * o4 + 0] = fmaf(x[i], w.x, y[4 * o4 + 0]);
            y[4 * o4 + 1] = fmaf(x[i], w.y, y[4 * o4 + 1]);
            y[4 * o4 + 2] = fmaf(x[i], w.z, y[4 * o4 + 2]);
            y[4 * o4 + 3] = fmaf(x[i], w.w, y[4 * o4 + 3]);
        }
    }
}
// y[o] = b[o] + sum_i W[o][i] * x[i]    torch Linear layout [NO][NI]; b may be null
template <int NI, int NO>
GLP_HD void linear(const float* W, const float* b, const float* x, float* y) {
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float a = b ? b[o] : 0.f;
#pragma unroll
        for (int i4 = 0; i4 < NI / 4; ++i4) {
            const F4 w = ld4(W + o * NI + 4 * i4);
            a = fmaf(w.x, x[4 * i4 + 0], a);
            a = fmaf(w.y, x[4 * i4 + 1], a);
            a = fmaf(w.z, x[4 * i4 + 2], a);
            a = fmaf(w.w, x[4 * i4 + 3], a);
        }
        y[o] = a;
    }
}

// f[27] -> h = embed(population features) [16], din = embedder_for_decoder([embed(exploration) ; embed(exploitation)]) [16]
GLP_HD void embed(const float* W, const float* f, float* h, float* din) {
    float cat[2 * E];
    mv_in<ND, E, false>(W + O_EMB, f, h);
    mv_in<ND, E, false>(W + O_EMB, f + ND, cat);
    mv_in<ND, E, false>(W + O_EMB, f + 2 * ND, cat + E);
    linear<2 * E, E>(W + O_DEMB, nullptr, cat, din);
}

// keys / values of one particle -> its rows of the lane's K / V tables; scaled queries stay with the thread
GLP_HD void project(const float* Wb, const float* qsrc, const float* kvsrc, float* qs, float* Krow, float* Vrow) {
    float t[E];
    mv_in<E, E, false>(Wb + BLK_WQ, qsrc, qs);
#pragma unroll
    for (int c = 0; c < E; ++c) qs[c] *= QSCALE;
    mv_in<E, E, false>(Wb + BLK_WK, kvsrc, t);
#pragma unroll
    for (int c = 0; c < E; c += 4) st4(Krow + c, t[c], t[c + 1], t[c + 2], t[c + 3]);
    mv_in<E, E, false>(Wb + BLK_WV, kvsrc, t);
#pragma unroll
    for (int c = 0; c < E; c += 4) st4(Vrow + c, t[c], t[c + 1], t[c + 2], t[c + 3]);
}

// 2^x for the softmax weights, x <= 0.  Device: one MUFU.EX2 (ex2.approx.ftz: 2 ulp, results below 2^-126 flush to zero -- such a
// weight is < 1e-38 of the row's largest one); exp2f without fast-math wraps the same instruction in a range fix-up (4 issue slots).
GLP_HD float exp2_neg(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}

// A/B knob GLP_PACKED_F32X2 (default off; `python -m metabox_b200.build --variant glp2 -DGLP_PACKED_F32X2`): the 4-term head dot product and
// the 4 accumulator updates go through Blackwell's packed FMUL2 / FFMA2 (two lanes per issue slot): 15 instead of 19 instructions per
// (key, head).  The dot product is then summed as (q0 k0 + q2 k2) + (q1 k1 + q3 k3); the host path mirrors that order exactly.
#if defined(GLP_PACKED_F32X2)
GLP_HD float score(const float* qs, const F4& k) {
#if defined(__CUDA_ARCH__)
    float2 p = __fmul2_rn(make_float2(qs[0], qs[1]), make_float2(k.x, k.y));
    p = __ffma2_rn(make_float2(qs[2], qs[3]), make_float2(k.z, k.w), p);
    return p.x + p.y;
#else
    return fmaf(qs[2], k.z, qs[0] * k.x) + fmaf(qs[3], k.w, qs[1] * k.y);
#endif
}
GLP_HD void accumulate(float* acc, float p, const F4& v) {
#if defined(__CUDA_ARCH__)
    const float2 pp = make_float2(p, p);
    const float2 a = __ffma2_rn(pp, make_float2(v.x, v.y), make_float2(acc[0], acc[1]));
    const float2 b = __ffma2_rn(pp, make_float2(v.z, v.w), make_float2(acc[2], acc[3]));
    acc[0] = a.x; acc[1] = a.y; acc[2] = b.x; acc[3] = b.y;
#else
    acc[0] = fmaf(p, v.x, acc[0]); acc[1] = fmaf(p, v.y, acc[1]); acc[2] = fmaf(p, v.z, acc[2]); acc[3] = fmaf(p, v.w, acc[3]);
#endif
}
#else
GLP_HD float score(const float* qs, const F4& k) { return fmaf(qs[3], k.w, fmaf(qs[2], k.z, fmaf(qs[1], k.y, qs[0] * k.x))); }
GLP_HD void accumulate(float* acc, float p, const F4& v) {
    acc[0] = fmaf(p, v.x, acc[0]); acc[1] = fmaf(p, v.y, acc[1]); acc[2] = fmaf(p, v.z, acc[2]); acc[3] = fmaf(p, v.w, acc[3]);
}
#endif

// softmax(q K^T / sqrt(4)) V for the four heads of one query; two passes over the NP keys (exact maximum first), then the output
// projection and the residual:  x = heads @ WO + res
GLP_HD void attend(const float* Wb, const float* K, const float* V, int NP, const float* qs, const float* res, float* x) {
    float m[NH], l[NH], acc[E];
#pragma unroll
    for (int h = 0; h < NH; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
#pragma unroll
    for (int c = 0; c < E; ++c) acc[c] = 0.f;
    for (int j = 0; j < NP; ++j) {
#pragma unroll
        for (int h = 0; h < NH; ++h) m[h] = fmaxf(m[h], score(qs + HK * h, ld4(K + j * E + HK * h)));
    }
    for (int j = 0; j < NP; ++j) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float p = exp2_neg(score(qs + HK * h, ld4(K + j * E + HK * h)) - m[h]);
            const F4 v = ld4(V + j * E + HK * h);
            l[h] += p;
            accumulate(acc + HK * h, p, v);
        }
    }
#pragma unroll
    for (int c = 0; c < E; ++c) { acc[c] = acc[c] / l[c / HK]; x[c] = res[c]; }
    mv_in<E, E, true>(Wb + BLK_WO, acc, x);
}

GLP_HD float sum16(const float* x) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < E; ++c) s += x[c];
    return s;
}
GLP_HD float sqdev16(const float* x, float mean) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < E; ++c) { const float d = x[c] - mean; s = fmaf(d, d, s); }
    return s;
}
// swarm norm (networks.py:68-72): statistics over all NP*16 entries of the lane, unbiased variance
GLP_HD float inv_std(float sqdev_total, int NP) { return 1.0f / sqrtf(sqdev_total / (float)(NP * E - 1) + NORM_EPS); }
GLP_HD void normalise(float* x, float mean, float inv) {
#pragma unroll
    for (int c = 0; c < E; ++c) x[c] = (x[c] - mean) * inv;
}
// y = F2(relu(F1 x)) + x   (in place)
GLP_HD void feed_forward(const float* Wb, float* x) {
    float t[E], y[E];
    linear<E, E>(Wb + BLK_F1W, Wb + BLK_F1B, x, t);
#pragma unroll
    for (int c = 0; c < E; ++c) t[c] = fmaxf(t[c], 0.f);
    linear<E, E>(Wb + BLK_F2W, Wb + BLK_F2B, t, y);
#pragma unroll
    for (int c = 0; c < E; ++c) x[c] += y[c];
}
// 16 -> 32 -> 8 -> 1 with LeakyReLU
GLP_HD float head(const float* Wh, const float* lg) {
    float a1[32], a2[8], o[1];
    linear<E, 32>(Wh + HD_L0W, Wh + HD_L0B, lg, a1);
#pragma unroll
    for (int c = 0; c < 32; ++c) a1[c] = a1[c] > 0.f ? a1[c] : LEAKY * a1[c];
    linear<32, 8>(Wh + HD_L1W, Wh + HD_L1B, a1, a2);
#pragma unroll
    for (int c = 0; c < 8; ++c) a2[c] = a2[c] > 0.f ? a2[c] : LEAKY * a2[c];
    linear<8, 1>(Wh + HD_L2W, Wh + HD_L2B, a2, o);
    return o[0];
}
GLP_HD void mu_sigma(const float* W, const float* lg, float& mu, float& sigma) {
    mu = (tanhf(head(W + O_MU, lg)) + 1.f) * 0.5f;
    sigma = (tanhf(head(W + O_SIGMA, lg)) + 1.f) * 0.5f * (MAX_SIGMA - MIN_SIGMA) + MIN_SIGMA;
}
GLP_HD float sample_action(float mu, float sigma, float eps) { return fminf(fmaxf(fmaf(sigma, eps, mu), 0.f), 1.f); }
}  // namespace glp
