// mbx_fdr.cuh -- the float32 screening pass of the FDR exemplar search of the RLEPSO generation (rlepso_optimizer.py:97-103), on its own so
// that the step kernel (mbx_rlepso.cu) and the loop micro-benchmark (scripts/ubench_fdr.cu) compile the SAME code.
#pragma once
#include "mbx_common.cuh"

// FDR pass 1: float32 screening (see the comment at the call site).  Work item = (pair of 2-rank blocks {m, nb-1-m},
// pair of dimensions): every item scans ~NP candidates.  Sign convention: a = pf[s]-pf[r] <= 0; the carried best is
// (nab, bb) = (-a_best, q_best); candidate wins iff  a*bb + nab*q < 0.
// Compile-time tuning knobs (defaults = the measured configuration; scripts/variant_sweep.sh builds and benches alternatives):
#ifndef MBX_FDR_UNROLL
#define MBX_FDR_UNROLL 2            // unroll factor of the FDR candidate scan
#endif
#define MBX_PRAGMA_(x) _Pragma(#x)
#define MBX_UNROLL(n) MBX_PRAGMA_(unroll n)
#define FDR_TAU 1e-6f
#define FDR_TINY 1e-35f
#define FDR_NEAR 2e-4f
#ifndef MBX_FDR_PRED
#define MBX_FDR_PRED 0
#endif
#ifndef MBX_FDR_AMBMUL
#define MBX_FDR_AMBMUL 0
#endif
#ifndef MBX_FDR_RCP
#define MBX_FDR_RCP 0
#endif
#ifndef MBX_FDR_SCALAR
#define MBX_FDR_SCALAR 0
#endif
#ifndef MBX_FDR_ABL
#define MBX_FDR_ABL 0      // diagnostic ablations for scripts/ubench_fdr.cu only (wrong results): 1 = no ambiguity test, 2 = no index, 4 = no replacement
#endif
#ifndef MBX_FDR_AHEAD
#define MBX_FDR_AHEAD 0
#endif
#if MBX_FDR_PRED
typedef float FdrIdx;      // sorted positions < 128 are exact in float32: the index rides the FMA pipe like the two other carried values
#else
typedef int FdrIdx;
#endif
struct FdrChain {          // two ranks x two dims
    float2 nab, bb;
    FdrIdx jx, jy;
    bool amb;
};
// MODE bit 0: -a under a predicate (FMA pipe), bit 1: q under a predicate, bit 2: the index under a predicate
template <int MODE>
__device__ __forceinline__ void fdr_take(float& nab, float& bb, float& jf, const float m, const float a, const float q, const float sf) {
    if (MODE == 7) {            // -a, q and the index (kept as a float) under a predicate
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %3, 0f00000000;\n\t@p sub.f32 %0, 0f80000000, %4;\n\t@p add.f32 %1, %5, 0f80000000;\n\t@p add.f32 %2, %6, 0f80000000;\n\t}"
            : "+f"(nab), "+f"(bb), "+f"(jf) : "f"(m), "f"(a), "f"(q), "f"(sf));
    } else if (MODE == 1) {     // -a only
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %3, 0f00000000;\n\t@p sub.f32 %0, 0f80000000, %4;\n\tselp.f32 %1, %5, %1, p;\n\tselp.f32 %2, %6, %2, p;\n\t}"
            : "+f"(nab), "+f"(bb), "+f"(jf) : "f"(m), "f"(a), "f"(q), "f"(sf));
    } else if (MODE == 3) {     // -a and q
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %3, 0f00000000;\n\t@p sub.f32 %0, 0f80000000, %4;\n\t@p add.f32 %1, %5, 0f80000000;\n\tselp.f32 %2, %6, %2, p;\n\t}"
            : "+f"(nab), "+f"(bb), "+f"(jf) : "f"(m), "f"(a), "f"(q), "f"(sf));
    } else if (MODE == 4) {     // the index only
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %3, 0f00000000;\n\tselp.f32 %0, %4, %0, p;\n\tselp.f32 %1, %5, %1, p;\n\t@p add.f32 %2, %6, 0f80000000;\n\t}"
            : "+f"(nab), "+f"(bb), "+f"(jf) : "f"(m), "f"(-a), "f"(q), "f"(sf));
    } else {                    // 5: -a and the index
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %3, 0f00000000;\n\t@p sub.f32 %0, 0f80000000, %4;\n\tselp.f32 %1, %5, %1, p;\n\t@p add.f32 %2, %6, 0f80000000;\n\t}"
            : "+f"(nab), "+f"(bb), "+f"(jf) : "f"(m), "f"(a), "f"(q), "f"(sf));
    }
}
struct FdrPrep { float2 a, q; };      // the part of a visit that does not depend on the carried state
__device__ __forceinline__ FdrPrep fdr_prep(const float2 s2, const float2 x2, const float2 npf, const float2 npx) {
#if MBX_FDR_SCALAR
    // scalar FADD / FMUL / FFMA issue every cycle on the FMA pipe; the packed f32x2 forms take 2 (FADD2, FMUL2) or 3 (FFMA2) cycles of a
    // resource they share with the ALU pipe (scripts/ubench_pipes.cu), which the compare / select half of this loop already saturates
    const float a = s2.x + npf.x;
    return FdrPrep{make_float2(a, a), make_float2(fabsf(x2.x + npx.x) + 1e-5f, fabsf(x2.y + npx.y) + 1e-5f)};
#else
    const float2 a = __fadd2_rn(s2, npf);
    const float2 dq = __fadd2_rn(x2, npx);
    return FdrPrep{a, make_float2(fabsf(dq.x) + 1e-5f, fabsf(dq.y) + 1e-5f)};
#endif
}
__device__ __forceinline__ void fdr_apply(FdrChain& c, const FdrPrep p, const FdrIdx s) {
    const float2 a = p.a, q = p.q;
#if MBX_FDR_SCALAR == 2
    // only two-operand scalar instructions on the FMA pipe (a three-register FFMA does not overlap with the ALU-pipe compares / selects,
    // FADD / FMUL do: scripts/ubench_pipes.cu): m = fl(a * bb) + fl(nab * q), two more roundings than the fused form, covered by FDR_TAU
    const float2 nr = make_float2(__fmul_rn(c.nab.x, q.x), __fmul_rn(c.nab.y, q.y));
    const float2 m = make_float2(__fadd_rn(__fmul_rn(a.x, c.bb.x), nr.x), __fadd_rn(__fmul_rn(a.y, c.bb.y), nr.y));
    const float2 t = make_float2(__fmul_rn(nr.x, FDR_TAU), __fmul_rn(nr.y, FDR_TAU));
    c.amb = c.amb || !(fabsf(m.x) > t.x) || !(fabsf(m.y) > t.y);
#else
#if MBX_FDR_SCALAR
    const float2 nr = make_float2(__fmul_rn(c.nab.x, q.x), __fmul_rn(c.nab.y, q.y));
    const float2 m = make_float2(__fmaf_rn(a.x, c.bb.x, nr.x), __fmaf_rn(a.y, c.bb.y, nr.y));
#else
    const float2 nr = __fmul2_rn(c.nab, q);
    const float2 m = __ffma2_rn(a, c.bb, nr);
#endif
#if MBX_FDR_ABL & 1
#elif MBX_FDR_AMBMUL
    // |m| <= TAU * nr  <=>  |m| / TAU <= nr (nr = nab * q >= 0): a packed multiply (2 cycles of the FMA pipe) instead of a packed FMA (3)
#if MBX_FDR_SCALAR
    const float2 ms = make_float2(__fmul_rn(m.x, 1.f / FDR_TAU), __fmul_rn(m.y, 1.f / FDR_TAU));
#else
    const float2 ms = __fmul2_rn(m, make_float2(1.f / FDR_TAU, 1.f / FDR_TAU));
#endif
    c.amb = c.amb || !(fabsf(ms.x) > nr.x) || !(fabsf(ms.y) > nr.y);
#else
#if MBX_FDR_SCALAR
    const float2 t = make_float2(__fmaf_rn(nr.x, FDR_TAU, FDR_TINY), __fmaf_rn(nr.y, FDR_TAU, FDR_TINY));
#else
    const float2 t = __ffma2_rn(nr, make_float2(FDR_TAU, FDR_TAU), make_float2(FDR_TINY, FDR_TINY));
#endif
    c.amb = c.amb || !(fabsf(m.x) > t.x) || !(fabsf(m.y) > t.y);
#endif
#endif
#if MBX_FDR_PRED
    // the three carried values of a chain are replaced under a predicate: ptxas turns `if` into FSEL / SEL, which all go to the ALU pipe
    // (one warp instruction per two cycles and scheduler); predicated FADD go to the FMA pipe, which has slack in this loop
    fdr_take<MBX_FDR_PRED>(c.nab.x, c.bb.x, c.jx, m.x, a.x, q.x, s);
    fdr_take<MBX_FDR_PRED>(c.nab.y, c.bb.y, c.jy, m.y, a.y, q.y, s);
#elif MBX_FDR_ABL & 4
    if (m.x < 0.f) c.jx += 1;
    if (m.y < 0.f) c.jy += 1;
#elif MBX_FDR_ABL & 2
    if (m.x < 0.f) { c.nab.x = -a.x; c.bb.x = q.x; }
    if (m.y < 0.f) { c.nab.y = -a.y; c.bb.y = q.y; }
#else
    if (m.x < 0.f) { c.nab.x = -a.x; c.bb.x = q.x; c.jx = s; }
    if (m.y < 0.f) { c.nab.y = -a.y; c.bb.y = q.y; c.jy = s; }
#endif
}
__device__ __forceinline__ void fdr_visit(FdrChain& c, const float2 s2, const float2 x2, const float2 npf, const float2 npx, const FdrIdx s) {
    fdr_apply(c, fdr_prep(s2, x2, npf, npx), s);
}
#if MBX_FDR_RCP
// PROTOTYPE (timing only): quotient keys a * rcp(q) with the candidate's sorted position in the 7 low mantissa bits; the carried state of a
// chain is (smallest key, second smallest key), both by FMNMX; the item is ambiguous when the two are within FDR_TAU_KEY of each other.
#define FDR_TAU_KEY 1.25e-4f
struct FdrKeys { float bx, by, sx, sy; };
__device__ __forceinline__ float2 fdr_keys(const float fs, const float2 x2, const float npf, const float2 npx, const unsigned s, const unsigned c7 = 0, const unsigned c128 = 0) {
    const float a = fs + npf;
#if MBX_FDR_SCALAR
    const float2 dq = make_float2(x2.x + npx.x, x2.y + npx.y);
#else
    const float2 dq = __fadd2_rn(x2, npx);
#endif
#if MBX_FDR_RCP >= 4
    // one reciprocal for the two dimensions: a / qx = (a / (qx qy)) * qy -- the MUFU pipe takes 8 cycles per warp instruction
    const float qx = fabsf(dq.x) + 1e-5f, qy = fabsf(dq.y) + 1e-5f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(qx, qy)));
    const float ar = __fmul_rn(a, r);
    const float2 ratio = make_float2(__fmul_rn(ar, qy), __fmul_rn(ar, qx));
#else
    float rx, ry;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(fabsf(dq.x) + 1e-5f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(fabsf(dq.y) + 1e-5f));
#if MBX_FDR_SCALAR
    const float2 ratio = make_float2(__fmul_rn(a, rx), __fmul_rn(a, ry));
#else
    const float2 ratio = __fmul2_rn(make_float2(a, a), make_float2(rx, ry));
#endif
#endif
#if MBX_FDR_RCP == 6
    // index bits by two integer multiply-adds (FMA pipe) instead of one LOP3 (ALU pipe): (bits >> 7) * 128 + s; c7 = 2^25 and c128 = 128 are
    // run-time values so that ptxas does not turn the multiplications back into shifts
    unsigned kx, ky;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(kx) : "r"(__float_as_uint(ratio.x)), "r"(c7));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(ky) : "r"(__float_as_uint(ratio.y)), "r"(c7));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(kx) : "r"(kx), "r"(c128), "r"(s));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ky) : "r"(ky), "r"(c128), "r"(s));
    return make_float2(__uint_as_float(kx), __uint_as_float(ky));
#else
    return make_float2(__uint_as_float((__float_as_uint(ratio.x) & 0xFFFFFF80u) | s), __uint_as_float((__float_as_uint(ratio.y) & 0xFFFFFF80u) | s));
#endif
}
__device__ __forceinline__ float fmin3(const float a, const float b, const float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ void fdr_push1(float& best, float& second, const float k) {
    second = fmaxf(best, fminf(k, second)); best = fminf(best, k);
}
// two keys at once: 5 instead of 6 min / max instructions
__device__ __forceinline__ void fdr_push2(float& best, float& second, const float k1, const float k2) {
#if MBX_FDR_RCP >= 2
    const float lo = fminf(k1, k2), hi = fmaxf(k1, k2);
    second = fmin3(second, hi, fmaxf(best, lo)); best = fminf(best, lo);
#else
    fdr_push1(best, second, k1); fdr_push1(best, second, k2);
#endif
}
__device__ __forceinline__ void fdr_visit_rcp(FdrKeys& c, const float fs, const float2 x2, const float npf, const float2 npx, const unsigned s) {
    const float2 k = fdr_keys(fs, x2, npf, npx, s);
    fdr_push1(c.bx, c.sx, k.x); fdr_push1(c.by, c.sy, k.y);
}
template <int D>
__device__ __forceinline__ void fdr_screen(const int NP, const int tid, const float* __restrict__ spx, const float2* __restrict__ spf2,
                                           const unsigned char* __restrict__ order, unsigned char* tgt, unsigned short* fbl, int* nfb) {
    constexpr int DP = D / 2;
    const int nb = (NP + 1) / 2, npair = (nb + 1) / 2;
    const float pf0 = spf2[0].x;
    const unsigned c128 = (unsigned)NP + 28u, c7 = c128 << 18;      // (timing prototype: 128 and 2^25 for NP = 100)
    for (int w = tid; w < npair * DP; w += MBX_NT) {
        const int pq = w / DP, d = 2 * (w - pq * DP);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int m = half ? nb - 1 - pq : pq;
            if (half && m == pq) break;
            const int ra = 2 * m, rb = ra + 1;
            const bool hasb = rb < NP;
            const float2 pxa = *reinterpret_cast<const float2*>(spx + ra * D + d);
            const float pfa = spf2[ra].x;
            const float2 pxb = hasb ? *reinterpret_cast<const float2*>(spx + rb * D + d) : pxa;
            const float pfb = hasb ? spf2[rb].x : pfa;
            const float2 npxa = make_float2(-pxa.x, -pxa.y), npxb = make_float2(-pxb.x, -pxb.y);
            FdrKeys A{__uint_as_float((unsigned)ra), __uint_as_float((unsigned)ra), INFINITY, INFINITY};
            FdrKeys B{__uint_as_float((unsigned)rb), __uint_as_float((unsigned)rb), INFINITY, INFINITY};
#if MBX_FDR_RCP == 3 || MBX_FDR_RCP == 5
            // software pipeline: the keys of candidates s + 2, s + 3 (loads, packed subtract, two reciprocals, packed multiply, index bits:
            // a ~60-cycle dependent sequence) are computed while the keys of s, s + 1 go through the min / max chains
            float2 ka1, ka2, kb1, kb2;
            {
                const float2 x2 = *reinterpret_cast<const float2*>(spx + d), y2 = *reinterpret_cast<const float2*>(spx + D + d);
                const float fs = spf2[0].x, gs = spf2[1].x;
                ka1 = fdr_keys(fs, x2, -pfa, npxa, 0u); ka2 = fdr_keys(gs, y2, -pfa, npxa, 1u);
                kb1 = fdr_keys(fs, x2, -pfb, npxb, 0u); kb2 = fdr_keys(gs, y2, -pfb, npxb, 1u);
            }
#pragma unroll 1
            for (int s = 0; s < ra; s += 2) {
                const int sn = min(s + 2, NP - 2);                              // the last look-ahead is never used
                const float2 x2 = *reinterpret_cast<const float2*>(spx + sn * D + d), y2 = *reinterpret_cast<const float2*>(spx + (sn + 1) * D + d);
                const float fs = spf2[sn].x, gs = spf2[sn + 1].x;
                const float2 na1 = fdr_keys(fs, x2, -pfa, npxa, (unsigned)sn), na2 = fdr_keys(gs, y2, -pfa, npxa, (unsigned)sn + 1);
                const float2 nb1 = fdr_keys(fs, x2, -pfb, npxb, (unsigned)sn), nb2 = fdr_keys(gs, y2, -pfb, npxb, (unsigned)sn + 1);
                fdr_push2(A.bx, A.sx, ka1.x, ka2.x); fdr_push2(A.by, A.sy, ka1.y, ka2.y);
                fdr_push2(B.bx, B.sx, kb1.x, kb2.x); fdr_push2(B.by, B.sy, kb1.y, kb2.y);
                ka1 = na1; ka2 = na2; kb1 = nb1; kb2 = nb2;
            }
#else
#pragma unroll 1
            for (int s = 0; s < ra; s += 2) {                                  // ra is even: candidates come in pairs
                const float2 x2 = *reinterpret_cast<const float2*>(spx + s * D + d), y2 = *reinterpret_cast<const float2*>(spx + (s + 1) * D + d);
                const float fs = spf2[s].x, gs = spf2[s + 1].x;
                const float2 ka1 = fdr_keys(fs, x2, -pfa, npxa, (unsigned)s, c7, c128), ka2 = fdr_keys(gs, y2, -pfa, npxa, (unsigned)s + 1, c7, c128);
                fdr_push2(A.bx, A.sx, ka1.x, ka2.x); fdr_push2(A.by, A.sy, ka1.y, ka2.y);
                const float2 kb1 = fdr_keys(fs, x2, -pfb, npxb, (unsigned)s, c7, c128), kb2 = fdr_keys(gs, y2, -pfb, npxb, (unsigned)s + 1, c7, c128);
                fdr_push2(B.bx, B.sx, kb1.x, kb2.x); fdr_push2(B.by, B.sy, kb1.y, kb2.y);
            }
#endif
            if (hasb) fdr_visit_rcp(B, pfa, pxa, -pfb, npxb, (unsigned)ra);
            int jax = __float_as_uint(A.bx) & 127, jay = __float_as_uint(A.by) & 127, jbx = __float_as_uint(B.bx) & 127, jby = __float_as_uint(B.by) & 127;
            bool amba = !(A.sx > A.bx * (1.f - FDR_TAU_KEY)) || !(A.sy > A.by * (1.f - FDR_TAU_KEY)) || !(pfa - pf0 < 1e30f);
            bool ambb = !(B.sx > B.bx * (1.f - FDR_TAU_KEY)) || !(B.sy > B.by * (1.f - FDR_TAU_KEY)) || !(pfb - pf0 < 1e30f);
            if (pfa == pf0) { jax = jay = 0; amba = false; }
            if (pfb == pf0) { jbx = jby = 0; ambb = false; }
            const int oa = order[ra] * D + d;
            tgt[oa] = (unsigned char)jax; tgt[oa + 1] = (unsigned char)jay;
            if (amba) { const int k = atomicAdd(nfb, 2); fbl[k] = (unsigned short)(ra * D + d); fbl[k + 1] = (unsigned short)(ra * D + d + 1); }
            if (hasb) {
                const int ob = order[rb] * D + d;
                tgt[ob] = (unsigned char)jbx; tgt[ob + 1] = (unsigned char)jby;
                if (ambb) { const int k = atomicAdd(nfb, 2); fbl[k] = (unsigned short)(rb * D + d); fbl[k + 1] = (unsigned short)(rb * D + d + 1); }
            }
        }
    }
}
#else
// G = dimension pairs per work item: an item is (pair of 2-rank blocks {m, nb-1-m}, G adjacent dimension pairs) = 4 G independent chains per
// thread.  D = 30 -> G = 3 and D = 20 -> G = 2 make 125 items for NP = 100: one round over the 128 threads instead of three, the loads of
// the candidate's cost and the loop bookkeeping are shared by G times more comparisons, and the scheduler sees G times more independent
// instructions per warp (the scan is bound by dependency and pipe-throttle stalls at 4 warps per scheduler, not by a saturated pipe).
#ifndef MBX_FDR_G30
#define MBX_FDR_G30 1
#endif
#ifndef MBX_FDR_G20
#define MBX_FDR_G20 1
#endif
template <int D> struct FdrGroup { static constexpr int G = (D == 30 && (D / 2) % MBX_FDR_G30 == 0) ? MBX_FDR_G30 : (D == 20 && (D / 2) % MBX_FDR_G20 == 0) ? MBX_FDR_G20 : 1; };
template <int D>
__device__ __forceinline__ void fdr_screen(const int NP, const int tid, const float* __restrict__ spx, const float2* __restrict__ spf2,
                                           const unsigned char* __restrict__ order, unsigned char* tgt, unsigned short* fbl, int* nfb) {
    static_assert(D % 2 == 0, "the packed FDR scan pairs adjacent dimensions");
    constexpr int G = FdrGroup<D>::G, DP = D / 2, DG = DP / G;
    const int nb = (NP + 1) / 2, npair = (nb + 1) / 2;
    const float pf0 = spf2[0].x;
    for (int w = tid; w < npair * DG; w += MBX_NT) {
        const int pq = w / DG, d = 2 * G * (w - pq * DG);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int m = half ? nb - 1 - pq : pq;
            if (half && m == pq) break;
            const int ra = 2 * m, rb = ra + 1;
            const bool hasb = rb < NP;
            const float2 pfa = spf2[ra], pfb = hasb ? spf2[rb] : pfa;
            const float2 npfa = make_float2(-pfa.x, -pfa.y), npfb = make_float2(-pfb.x, -pfb.y);
            float2 pxa[G], npxa[G], npxb[G];
            FdrChain A[G], B[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                pxa[g] = *reinterpret_cast<const float2*>(spx + ra * D + d + 2 * g);
                const float2 pxb = hasb ? *reinterpret_cast<const float2*>(spx + rb * D + d + 2 * g) : pxa[g];
                npxa[g] = make_float2(-pxa[g].x, -pxa[g].y); npxb[g] = make_float2(-pxb.x, -pxb.y);
                A[g] = FdrChain{make_float2(0.f, 0.f), make_float2(1e-5f, 1e-5f), (FdrIdx)ra, (FdrIdx)ra, false};      // start from j = i (ratio 0)
                B[g] = FdrChain{make_float2(0.f, 0.f), make_float2(1e-5f, 1e-5f), (FdrIdx)rb, (FdrIdx)rb, false};
            }
            FdrIdx si = 0;
#if MBX_FDR_AHEAD
            // the state-independent half of candidate s + 1 (loads, two packed subtractions, |.| + 1e-5) is issued before the dependent half of
            // candidate s: the loop-carried chain (FMUL2 -> FFMA2 -> FSETP -> select) never waits for a shared-memory load
            FdrPrep pa[G], pb[G];
            {
                const float2 s2 = spf2[0];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 x2 = *reinterpret_cast<const float2*>(spx + d + 2 * g);
                    pa[g] = fdr_prep(s2, x2, npfa, npxa[g]); pb[g] = fdr_prep(s2, x2, npfb, npxb[g]);
                }
            }
            MBX_UNROLL(MBX_FDR_UNROLL)
            for (int s = 0; s < ra; ++s, si += 1) {
                const float2 s2 = spf2[s + 1];                                  // s + 1 <= ra < NP
                FdrPrep na[G], nb_[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 x2 = *reinterpret_cast<const float2*>(spx + (s + 1) * D + d + 2 * g);
                    na[g] = fdr_prep(s2, x2, npfa, npxa[g]); nb_[g] = fdr_prep(s2, x2, npfb, npxb[g]);
                }
#pragma unroll
                for (int g = 0; g < G; ++g) { fdr_apply(A[g], pa[g], si); fdr_apply(B[g], pb[g], si); pa[g] = na[g]; pb[g] = nb_[g]; }
            }
#else
            MBX_UNROLL(MBX_FDR_UNROLL)
            for (int s = 0; s < ra; ++s, si += 1) {
                const float2 s2 = spf2[s];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 x2 = *reinterpret_cast<const float2*>(spx + s * D + d + 2 * g);
                    fdr_visit(A[g], s2, x2, npfa, npxa[g], si);
                    fdr_visit(B[g], s2, x2, npfb, npxb[g], si);
                }
            }
#endif
            // best-cost group (pf[r] == pf[0]): every candidate ties with j = i at quotient 0, np.argmin takes the lowest
            // original index = sorted position 0 (stable sort).  Plateaus (Gallagher, step functions) make this group large.
            const int oa = order[ra] * D + d, ob = hasb ? order[rb] * D + d : 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (hasb) fdr_visit(B[g], pfa, pxa[g], npfb, npxb[g], (FdrIdx)ra);      // candidate s = ra for the block's second rank
                if (pfa.x == pf0) { A[g].jx = A[g].jy = 0; A[g].amb = false; }
                if (pfb.x == pf0) { B[g].jx = B[g].jy = 0; B[g].amb = false; }
                tgt[oa + 2 * g] = (unsigned char)A[g].jx; tgt[oa + 2 * g + 1] = (unsigned char)A[g].jy;
                if (A[g].amb) { const int k = atomicAdd(nfb, 2); fbl[k] = (unsigned short)(ra * D + d + 2 * g); fbl[k + 1] = (unsigned short)(ra * D + d + 2 * g + 1); }
                if (hasb) {
                    tgt[ob + 2 * g] = (unsigned char)B[g].jx; tgt[ob + 2 * g + 1] = (unsigned char)B[g].jy;
                    if (B[g].amb) { const int k = atomicAdd(nfb, 2); fbl[k] = (unsigned short)(rb * D + d + 2 * g); fbl[k + 1] = (unsigned short)(rb * D + d + 2 * g + 1); }
                }
            }
        }
    }
}
#endif
