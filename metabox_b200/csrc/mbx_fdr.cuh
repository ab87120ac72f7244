// mbx_fdr.cuh -- the float32 screening pass of the FDR exemplar search of the RLEPSO generation (rlepso_optimizer.py:97-103), on its own so
// that the step kernel (mbx_rlepso.cu) and the loop micro-benchmark (scripts/ubench_fdr.cu) compile the SAME code.
//
// For every (particle r, dimension d) the reference takes  argmin_j (pf[j] - pf[r]) / (|px[j,d] - px[r,d]| + 1e-5)  in float64, first index on
// ties.  j = r gives 0, so only particles with a strictly better pbest -- the ones BEFORE r in the stable cost order -- can win.  The screening
// pass finds the float32 winner of every (r, d) item and marks the items whose decision is within rounding ("ambiguous"); those, and only
// those, are redone with the reference's own float64 divisions by pass 2 (mbx_rlepso.cu).  Work item = (pair of 2-rank blocks {m, nb-1-m},
// pair of adjacent dimensions): 4 chains per thread, ~NP candidates per item whatever m is.
//
// Two formulations (profiles/r2z_ubench_fdr.txt has the timings of these and of the variants that were tried and dropped):
//   MBX_FDR_RCP = 1 (default)  quotient KEYS  a * rcp(q)  with the candidate's sorted position in the 7 low mantissa bits; a chain carries its
//                              smallest and second smallest key (FMNMX / FMNMX3), nothing is selected, no predicate is written
//   MBX_FDR_RCP = 0            cross-multiplication  a q' < a' q  with the carried best (a', q', index) replaced by selects (round 1)
#pragma once
#include "mbx_common.cuh"

#ifndef MBX_FDR_RCP
#define MBX_FDR_RCP 1
#endif
#ifndef MBX_FDR_UNROLL
#define MBX_FDR_UNROLL 2            // unroll factor of the cross-multiplication scan
#endif
#define MBX_PRAGMA_(x) _Pragma(#x)
#define MBX_UNROLL(n) MBX_PRAGMA_(unroll n)
#define FDR_NEAR 2e-4f              // pass 2 re-divides the candidates whose float32 quotient is within FDR_NEAR of the screening winner's
#define FDR_TINY 1e-35f

// the sorted costs live in the float2 array `spf2` of the step kernel: as (f, f) pairs for the packed cross-multiplication scan, as plain
// consecutive floats for the key scan (two candidates' costs in one 8-byte load)
#if MBX_FDR_RCP
#define FDR_COST(spf2, r) (reinterpret_cast<const float*>(spf2)[r])
__device__ __forceinline__ void fdr_store_cost(float2* spf2, const int r, const float f) { reinterpret_cast<float*>(spf2)[r] = f; }
#else
#define FDR_COST(spf2, r) ((spf2)[r].x)
__device__ __forceinline__ void fdr_store_cost(float2* spf2, const int r, const float f) { spf2[r] = make_float2(f, f); }
#endif
// flag the two items (r, d), (r, d + 1) for pass 2
__device__ __forceinline__ void fdr_flag(unsigned short* fbl, int* nfb, const int item) {
    const int k = atomicAdd(nfb, 2);
    fbl[k] = (unsigned short)item; fbl[k + 1] = (unsigned short)(item + 1);
}

#define FDR_NO_WINNER 255          // screening result of an item without a usable float32 winner: pass 2 divides for every candidate
#if MBX_FDR_RCP
// ---- reciprocal keys ---------------------------------------------------------------------------------------------------------------------
// key(s) = trunc7(fl(a_s * rcp(q_s))) | s,   a_s = fl(pf[s] - pf[r]) <= 0,   q_s = fl(|fl(px[s,d] - px[r,d])| + 1e-5f) > 0.
// Error budget against the float64 quotient rho_s of the same float32 inputs: a_s 2^-24, q_s 2^-23 (difference and sum, both terms >= 0) and
// 2^-25 (1e-5f vs 1e-5), rcp.approx 2^-23 (PTX: at most 1 ulp), the product 2^-24, the 7 replaced mantissa bits 2^-16:
//     key(s) = rho_s (1 + e),  |e| < DELTA = 2^-15.9.
// Let b be the smallest key (index j) and c the second smallest.  If c > b (1 - FDR_TAU_KEY) then for every s != j
//     |rho_s| <= |c| / (1 - DELTA) < |b| (1 - TAU) / (1 - DELTA) <= |b| / (1 + DELTA) <= |rho_j|      (TAU >= 2 DELTA / (1 + DELTA) = 2^-14.9),
// i.e. j is the float64 argmin and it is unique, so the tie rule never applies.  Otherwise the item goes to pass 2, whose candidate filter
// (FDR_NEAR = 2e-4 around the screening winner) covers the 2 DELTA = 3.3e-5 the true minimum can be away from it.  Items the bound does not
// cover are flagged outright: a cost span that is not finite or could overflow the quotient, and winners so small that the product may
// have lost bits to underflow.  Keys of equal-cost candidates (a_s = 0) are +denormals (the index bits): they lose against every negative
// key, and a chain without a negative key belongs to the best-cost group, which the tie rule sends to sorted position 0 (below).
#define FDR_TAU_KEY 6.103515625e-5f          // 2^-14
struct FdrKeys { float bx, by, sx, sy; };    // smallest / second smallest key of the two dimensions of a chain pair
__device__ __forceinline__ float2 fdr_keys(const float fs, const float2 x2, const float npf, const float2 npx, const unsigned s) {
    const float a = fs + npf;
    const float2 dq = __fadd2_rn(x2, npx);
    float rx, ry;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(fabsf(dq.x) + 1e-5f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(fabsf(dq.y) + 1e-5f));
    const float2 ratio = __fmul2_rn(make_float2(a, a), make_float2(rx, ry));
    return make_float2(__uint_as_float((__float_as_uint(ratio.x) & 0xFFFFFF80u) | s), __uint_as_float((__float_as_uint(ratio.y) & 0xFFFFFF80u) | s));
}
__device__ __forceinline__ float fdr_min3(const float a, const float b, const float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;          // FMNMX3 (sm_100)
}
__device__ __forceinline__ void fdr_push1(float& best, float& second, const float k) {
    second = fmaxf(best, fminf(k, second)); best = fminf(best, k);
}
__device__ __forceinline__ void fdr_push2(float& best, float& second, const float k1, const float k2) {   // two keys: 5 instead of 6 min / max
    const float lo = fminf(k1, k2), hi = fmaxf(k1, k2);
    second = fdr_min3(second, hi, fmaxf(best, lo)); best = fminf(best, lo);
}
// winner and ambiguity of one finished chain.  `wide` (the cost span is outside the range of the bound) and winners below 1e-30 in magnitude
// (the product may have underflowed) have no usable screening winner: FDR_NO_WINNER tells pass 2 to divide for EVERY candidate of the item.
__device__ __forceinline__ int fdr_pick(const float best, const float second, const bool wide, bool& amb) {
    const bool lost = wide || !(best < -1e-30f);
    amb = lost || !(second > best * (1.f - FDR_TAU_KEY));
    return lost ? FDR_NO_WINNER : (int)(__float_as_uint(best) & 127u);
}
template <int D>
__device__ __forceinline__ void fdr_screen(const int NP, const int tid, const float* __restrict__ spx, const float2* __restrict__ spf2,
                                           const unsigned char* __restrict__ order, unsigned char* tgt, unsigned short* fbl, int* nfb) {
    static_assert(D % 2 == 0, "the FDR scan pairs adjacent dimensions");
    constexpr int DP = D / 2;
    const float* __restrict__ spf = reinterpret_cast<const float*>(spf2);
    const int nb = (NP + 1) / 2, npair = (nb + 1) / 2;
    const float pf0 = spf[0];
    const int row1 = min(1, NP - 1);
    for (int w = tid; w < npair * DP; w += MBX_NT) {
        const int pq = w / DP, d = 2 * (w - pq * DP);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int m = half ? nb - 1 - pq : pq;
            if (half && m == pq) break;
            const int ra = 2 * m, rb = ra + 1;
            const bool hasb = rb < NP;
            const float2 pxa = *reinterpret_cast<const float2*>(spx + ra * D + d);
            const float pfa = spf[ra];
            const float2 pxb = hasb ? *reinterpret_cast<const float2*>(spx + rb * D + d) : pxa;
            const float pfb = hasb ? spf[rb] : pfa;
            const float2 npxa = make_float2(-pxa.x, -pxa.y), npxb = make_float2(-pxb.x, -pxb.y);
            // j = i: quotient 0, key = the position bits (a +denormal)
            FdrKeys A{__uint_as_float((unsigned)ra), __uint_as_float((unsigned)ra), INFINITY, INFINITY};
            FdrKeys B{__uint_as_float((unsigned)rb), __uint_as_float((unsigned)rb), INFINITY, INFINITY};
            // candidates come in pairs (ra is even); software pipeline: the keys of s + 2, s + 3 (loads, subtractions, reciprocals, product,
            // position bits: a long dependent sequence) are computed while the keys of s, s + 1 go through the min / max chains
            float2 ka1, ka2, kb1, kb2;
            {
                const float2 x2 = *reinterpret_cast<const float2*>(spx + d), y2 = *reinterpret_cast<const float2*>(spx + row1 * D + d);
                const float fs = spf[0], gs = spf[row1];
                ka1 = fdr_keys(fs, x2, -pfa, npxa, 0u); ka2 = fdr_keys(gs, y2, -pfa, npxa, 1u);
                kb1 = fdr_keys(fs, x2, -pfb, npxb, 0u); kb2 = fdr_keys(gs, y2, -pfb, npxb, 1u);
            }
            const float* px = spx + 2 * D + d;                                  // rows s + 2, s + 3
            const float* pf = spf + 2;
#pragma unroll 1
            for (unsigned s = 2; s <= (unsigned)ra; s += 2, px += 2 * D, pf += 2) {
                // the look-ahead of the last trip (rows ra, ra + 1 <= NP: at most one row past the sorted copy, still inside the CTA's region
                // A / cost array) is computed and dropped
                const float2 x2 = *reinterpret_cast<const float2*>(px), y2 = *reinterpret_cast<const float2*>(px + D);
                const float2 fg = *reinterpret_cast<const float2*>(pf);
                const float2 na1 = fdr_keys(fg.x, x2, -pfa, npxa, s), na2 = fdr_keys(fg.y, y2, -pfa, npxa, s + 1);
                const float2 nb1 = fdr_keys(fg.x, x2, -pfb, npxb, s), nb2 = fdr_keys(fg.y, y2, -pfb, npxb, s + 1);
                fdr_push2(A.bx, A.sx, ka1.x, ka2.x); fdr_push2(A.by, A.sy, ka1.y, ka2.y);
                fdr_push2(B.bx, B.sx, kb1.x, kb2.x); fdr_push2(B.by, B.sy, kb1.y, kb2.y);
                ka1 = na1; ka2 = na2; kb1 = nb1; kb2 = nb2;
            }
            if (hasb) {                                                        // candidate s = ra for the block's second rank
                const float2 k = fdr_keys(pfa, pxa, -pfb, npxb, (unsigned)ra);
                fdr_push1(B.bx, B.sx, k.x); fdr_push1(B.by, B.sy, k.y);
            }
            // the bound above needs a finite cost span whose quotients cannot overflow (a >= -1e30, rcp(q) <= 1e5)
            const bool widea = !(pfa - pf0 < 1e30f), wideb = !(pfb - pf0 < 1e30f);
            bool ambax, ambay, ambbx, ambby;
            int jax = fdr_pick(A.bx, A.sx, widea, ambax), jay = fdr_pick(A.by, A.sy, widea, ambay);
            int jbx = fdr_pick(B.bx, B.sx, wideb, ambbx), jby = fdr_pick(B.by, B.sy, wideb, ambby);
            // best-cost group (pf[r] == pf[0]): every candidate ties with j = i at quotient 0, np.argmin takes the lowest
            // original index = sorted position 0 (stable sort).  Plateaus (Gallagher, step functions) make this group large.
            if (pfa == pf0) { jax = jay = 0; ambax = ambay = false; }
            if (pfb == pf0) { jbx = jby = 0; ambbx = ambby = false; }
            const int oa = order[ra] * D + d;
            tgt[oa] = (unsigned char)jax; tgt[oa + 1] = (unsigned char)jay;
            if (ambax || ambay) fdr_flag(fbl, nfb, ra * D + d);
            if (hasb) {
                const int ob = order[rb] * D + d;
                tgt[ob] = (unsigned char)jbx; tgt[ob + 1] = (unsigned char)jby;
                if (ambbx || ambby) fdr_flag(fbl, nfb, rb * D + d);
            }
        }
    }
}

#else
// ---- cross-multiplication (round 1; kept as the A/B reference) ------------------------------------------------------------------------
// Sign convention: a = pf[s] - pf[r] <= 0; the carried best is (nab, bb) = (-a_best, q_best); a candidate wins iff  a * bb + nab * q < 0,
// packed f32x2 over two adjacent dimensions, two ranks per thread.  Every comparison whose margin is below FDR_TAU (>= 3x the worst-case
// float32 rounding of both sides), or is NaN, marks the item ambiguous.
#define FDR_TAU 1e-6f
struct FdrChain {          // one rank x two dims
    float2 nab, bb;
    int jx, jy;
    bool amb;
};
__device__ __forceinline__ void fdr_visit(FdrChain& c, const float2 s2, const float2 x2, const float2 npf, const float2 npx, const int s) {
    const float2 a = __fadd2_rn(s2, npf);
    const float2 dq = __fadd2_rn(x2, npx);
    const float2 q = make_float2(fabsf(dq.x) + 1e-5f, fabsf(dq.y) + 1e-5f);
    const float2 nr = __fmul2_rn(c.nab, q);
    const float2 m = __ffma2_rn(a, c.bb, nr);
    const float2 t = __ffma2_rn(nr, make_float2(FDR_TAU, FDR_TAU), make_float2(FDR_TINY, FDR_TINY));
    c.amb = c.amb || !(fabsf(m.x) > t.x) || !(fabsf(m.y) > t.y);
    if (m.x < 0.f) { c.nab.x = -a.x; c.bb.x = q.x; c.jx = s; }
    if (m.y < 0.f) { c.nab.y = -a.y; c.bb.y = q.y; c.jy = s; }
}
template <int D>
__device__ __forceinline__ void fdr_screen(const int NP, const int tid, const float* __restrict__ spx, const float2* __restrict__ spf2,
                                           const unsigned char* __restrict__ order, unsigned char* tgt, unsigned short* fbl, int* nfb) {
    static_assert(D % 2 == 0, "the packed FDR scan pairs adjacent dimensions");
    constexpr int DP = D / 2;
    const int nb = (NP + 1) / 2, npair = (nb + 1) / 2;
    const float pf0 = spf2[0].x;
    for (int w = tid; w < npair * DP; w += MBX_NT) {
        const int pq = w / DP, d = 2 * (w - pq * DP);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int m = half ? nb - 1 - pq : pq;
            if (half && m == pq) break;
            const int ra = 2 * m, rb = ra + 1;
            const bool hasb = rb < NP;
            const float2 pxa = *reinterpret_cast<const float2*>(spx + ra * D + d), pfa = spf2[ra];
            const float2 pxb = hasb ? *reinterpret_cast<const float2*>(spx + rb * D + d) : pxa, pfb = hasb ? spf2[rb] : pfa;
            const float2 npxa = make_float2(-pxa.x, -pxa.y), npfa = make_float2(-pfa.x, -pfa.y);
            const float2 npxb = make_float2(-pxb.x, -pxb.y), npfb = make_float2(-pfb.x, -pfb.y);
            FdrChain A{make_float2(0.f, 0.f), make_float2(1e-5f, 1e-5f), ra, ra, false};      // start from j = i (ratio 0)
            FdrChain B{make_float2(0.f, 0.f), make_float2(1e-5f, 1e-5f), rb, rb, false};
            MBX_UNROLL(MBX_FDR_UNROLL)
            for (int s = 0; s < ra; ++s) {
                const float2 x2 = *reinterpret_cast<const float2*>(spx + s * D + d), s2 = spf2[s];
                fdr_visit(A, s2, x2, npfa, npxa, s);
                fdr_visit(B, s2, x2, npfb, npxb, s);
            }
            if (hasb) fdr_visit(B, pfa, pxa, npfb, npxb, ra);                  // candidate s = ra for the block's second rank
            // best-cost group (pf[r] == pf[0]): every candidate ties with j = i at quotient 0, np.argmin takes the lowest
            // original index = sorted position 0 (stable sort).  Plateaus (Gallagher, step functions) make this group large.
            if (pfa.x == pf0) { A.jx = A.jy = 0; A.amb = false; }
            if (pfb.x == pf0) { B.jx = B.jy = 0; B.amb = false; }
            const int oa = order[ra] * D + d;
            tgt[oa] = (unsigned char)A.jx; tgt[oa + 1] = (unsigned char)A.jy;
            if (A.amb) fdr_flag(fbl, nfb, ra * D + d);
            if (hasb) {
                const int ob = order[rb] * D + d;
                tgt[ob] = (unsigned char)B.jx; tgt[ob + 1] = (unsigned char)B.jy;
                if (B.amb) fdr_flag(fbl, nfb, rb * D + d);
            }
        }
    }
}
#endif
