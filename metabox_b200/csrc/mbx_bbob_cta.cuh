// mbx_bbob_cta.cuh -- CTA-cooperative evaluation of a tile of BBOB candidates.
//
// The 24 BBOB shapes (reference problem/bbob.py:149-890), organised so that
// every thread of the CTA works whatever the number of candidates n (n = NP for the population kernels, n = 1 for the
// DE-DDQN single trial):
//   * element-wise transforms (T_osz, T_asy, cos / Weierstrass / Katsuura series, pair terms ...) run over all (i, k)
//     pairs of the [n, D] tile, one pair per thread per round -- 100 candidates use 128 of 128 lanes instead of 100, and a
//     single candidate is spread over D threads instead of one;
//   * the D x D rotations are register-tiled (TI candidates x TK rows per thread, operands from shared memory);
//   * the reductions over k stay sequential per candidate (thread i sums row i in ascending k).
// Scratch: B[n][ZS] float64 in shared memory, ZS = D + 1 (odd: rows hit distinct banks for the per-candidate passes).
// ALL threads of the CTA must call (barriers inside).  Returns f(x) incl. bias and boundary handling for candidate
// `MBX_TID` (valid when MBX_TID < n and the candidate is not masked out); requires n <= MBX_BDIM.
#pragma once
#include "mbx_bbob.cuh"

template <int D>
struct BbobCta {
    static constexpr int ZS = D + 1;
    static constexpr int TK = (D % 6 == 0) ? 6 : ((D % 5 == 0) ? 5 : ((D % 4 == 0) ? 4 : ((D % 3 == 0) ? 3 : ((D % 2 == 0) ? 2 : 1))));
    static constexpr int TI = (D >= 20) ? 4 : 2;
    static constexpr int KT = D / TK;
    static __host__ __device__ constexpr int scratch_doubles(int n) { return n * ZS; }
};

// B[i][:] <- M B[i][:] (+ add), rows i < n.  Register tile TI x TK; all k-tiles of a row live in the same round, so the
// in-place update is safe with one barrier between the reads and the writes of a round.
template <int D, int TI, int TK>
__device__ __forceinline__ void matvec_cta(const double* __restrict__ M, double* B, const int n, const double add) {
    constexpr int ZS = BbobCta<D>::ZS, KT = D / TK;
    const int tid = MBX_TID, NT = MBX_BDIM;
    const int IT = (n + TI - 1) / TI, per_round = NT / KT;
    const int lit = tid / KT, kt = tid - lit * KT;
    for (int it0 = 0; it0 < IT; it0 += per_round) {
        const int it = it0 + lit;
        const bool on = lit < per_round && it < IT;
        double acc[TI][TK];
#pragma unroll
        for (int j = 0; j < TI; ++j)
#pragma unroll
            for (int m = 0; m < TK; ++m) acc[j][m] = 0.;
        if (on) {
            const double* yr[TI];
#pragma unroll
            for (int j = 0; j < TI; ++j) yr[j] = B + min(it * TI + j, n - 1) * ZS;
            const double* mr = M + kt * TK * D;
            // one output per thread (single candidate; M may live in global memory): all D operand loads in flight at once
            constexpr int UNR = (TI * TK == 1) ? D : 2;
#pragma unroll UNR
            for (int d = 0; d < D; ++d) {
                double yv[TI], mv[TK];
#pragma unroll
                for (int j = 0; j < TI; ++j) yv[j] = yr[j][d];
#pragma unroll
                for (int m = 0; m < TK; ++m) mv[m] = mr[m * D + d];
#pragma unroll
                for (int j = 0; j < TI; ++j)
#pragma unroll
                    for (int m = 0; m < TK; ++m) acc[j][m] = fma(mv[m], yv[j], acc[j][m]);
            }
        }
        MBX_SYNC();
        if (on) {
#pragma unroll
            for (int j = 0; j < TI; ++j)
                if (it * TI + j < n) {
#pragma unroll
                    for (int m = 0; m < TK; ++m) B[(it * TI + j) * ZS + kt * TK + m] = acc[j][m] + add;
                }
        }
        MBX_SYNC();
    }
}
template <int D>
__device__ __forceinline__ void matvec_cta_auto(const double* __restrict__ M, double* B, const int n, const double add = 0.) {
    if (n * BbobCta<D>::KT >= 64) matvec_cta<D, BbobCta<D>::TI, BbobCta<D>::TK>(M, B, n, add);
    else matvec_cta<D, 1, 1>(M, B, n, add);                        // few candidates: one output per thread
}

// Weierstrass (bbob.py:620-625): sum_{j<12} 0.5^j cos(2 pi 3^j (z + 0.5)).  (cos, sin) of the base angle once (sincospi:
// exact range reduction), then the triple-angle map (c, s) -> (c (1 - 4 s^2), s (4 c^2 - 1)) eleven times.  The pair form has
// no 1/sin blow-up: an angle error triples per step, so term j carries <= 3^j ulp * 0.5^j weight (~1e-14 in total), below the
// reference's own cos(fl(2 pi (z + 0.5)) 3^j) argument rounding.
__device__ __forceinline__ double weierstrass_inner(const double z) {
    double sn, cs;
    sincospi(2. * (z + 0.5), &sn, &cs);
    double inner = cs, a = 0.5;
#pragma unroll
    for (int j = 1; j < 12; ++j) {
        const double c2 = cs * cs, s2 = sn * sn;
        cs = cs * (1. - 4. * s2);
        sn = sn * (4. * c2 - 1.);
        inner = fma(a, cs, inner);
        a *= 0.5;
    }
    return inner;
}
// Katsuura (bbob.py:855-866): (1 + (k+1) sum_{j=1..32} |2^j z - round(2^j z)| / 2^j)^expo.  Every term of the sum is exact in
// binary64 (power-of-two scalings, |2^j z| < 2^51); round() through the 1.5 * 2^52 magic constant (ties give the same
// |.| = 0.5 as the reference's floor(t + .5)).
__device__ __forceinline__ double katsuura_term(const double z, const int k, const double expo) {
    double acc = 0.;
#pragma unroll
    for (int j = 1; j <= 32; ++j) {
        const double p2 = (double)(1ull << j), ip2 = 1. / (double)(1ull << j);
        const double t = p2 * z;
        const double r = (t + 6755399441055744.) - 6755399441055744.;
        acc = fma(fabs(t - r), ip2, acc);
    }
    return exp(expo * log(1. + (double)(k + 1) * acc));
}

// merge of two partial Gallagher scans (disjoint peak subsets)
__device__ __forceinline__ void gal_merge(GalScan& a, const float s1, const float lo1, const float hi2, const int p1) {
    const float ahi = a.s1 + (a.s1 - a.lo1), bhi = s1 + (s1 - lo1);
    if (s1 > a.s1 || (s1 == a.s1 && p1 < a.p1)) { a.hi2 = fmaxf(fmaxf(a.hi2, hi2), ahi); a.s1 = s1; a.lo1 = lo1; a.p1 = p1; }
    else a.hi2 = fmaxf(fmaxf(a.hi2, hi2), bhi);
}

// Gallagher max_p w_p exp(coef acc_p) for the rows of B (= R x); float32 screening + exact float64 winner scheme as
// described at GalScan in mbx_bbob.cuh.  Each candidate is scanned by GP threads (GP = 1 for a full population,
// up to 32 for a single candidate), every thread taking the peaks p = g (mod GP); partial scans are merged by shuffles.
// `tile`: optional staging buffer for the float32 peak table (else __ldg).  Result valid in thread i (< n).
template <int D>
__device__ __forceinline__ double gallagher_cta(const double* __restrict__ G, const int P, const double* B, const int n,
                                                const unsigned char* only_mask, float4* tile, const int tile_bytes) {
    constexpr int HD = D / 2, ZS = BbobCta<D>::ZS;
    PROF_INIT
    const int tid = MBX_TID, NT = MBX_BDIM;
    const double* gC = G + ConstView<D>::kDoubles + P * D;
    const double* gw = gC + P * D;
    const double* gRy = gw + P;
    const int o32 = (ConstView<D>::kDoubles + 3 * P * D + P + 1) & ~1;
    const float4* g32 = reinterpret_cast<const float4*>(G + o32);
    const float* lnw = reinterpret_cast<const float*>(G + o32 + P * D);
    const double coef = -0.5 / (double)D;
    const float coef32 = (float)coef;
    int GP = 1;
    while (GP < 32 && GP * 2 * n <= NT) GP *= 2;
    const int ci = tid / GP, g = tid - ci * GP;                        // candidate, peak residue
    const bool act = ci < n && !(only_mask && !only_mask[ci]);
    float2 yf[HD];
    float ymax = 0.f;
#pragma unroll
    for (int k = 0; k < HD; ++k) {
        yf[k] = act ? make_float2((float)B[ci * ZS + 2 * k], (float)B[ci * ZS + 2 * k + 1]) : make_float2(0.f, 0.f);
        ymax = fmaxf(ymax, fmaxf(fabsf(yf[k].x), fabsf(yf[k].y)));
    }
    const float K1 = -coef32 * 2.f * 1.3e-7f * (ymax + 5.f * sqrtf((float)D) + 1.f) * sqrtf(31.7f * (float)D);
    GalScan sc;
    sc.p1 = 0x7fffffff;
    auto visit = [&](const float4* row, const float lw, const int p) {
        float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < HD; ++k) {
            const float4 c4 = row[k];
            const float2 z = __fadd2_rn(yf[k], make_float2(-c4.z, -c4.w));
            acc2 = __ffma2_rn(__fmul2_rn(make_float2(c4.x, c4.y), z), z, acc2);
        }
        sc.visit(acc2.x + acc2.y, lw, coef32, K1, p);
    };
    int TP = tile != nullptr ? tile_bytes / (16 * HD + 4) : 0;     // peaks per staged tile; a tile too small for one peak is not used
    if (TP * HD > 2 * (int)MBX_BDIM) TP = 2 * (int)MBX_BDIM / HD;                // two prefetch registers per thread
    if (TP < 1) {
        if (act) {
#pragma unroll 1
            for (int p = g; p < P; p += GP) {
                float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < HD; ++k) {
                    const float4 c4 = __ldg(g32 + p * HD + k);
                    const float2 z = __fadd2_rn(yf[k], make_float2(-c4.z, -c4.w));
                    acc2 = __ffma2_rn(__fmul2_rn(make_float2(c4.x, c4.y), z), z, acc2);
                }
                sc.visit(acc2.x + acc2.y, __ldg(lnw + p), coef32, K1, p);
            }
        }
    } else {
        float* tl = reinterpret_cast<float*>(tile + TP * HD);
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        float rl = 0.f;
        auto fetch = [&](const int p0) {
            const int m = min(TP, P - p0);
            if (tid < m * HD) r0 = __ldg(g32 + p0 * HD + tid);
            if (tid + NT < m * HD) r1 = __ldg(g32 + p0 * HD + tid + NT);
            if (tid < m) rl = __ldg(lnw + p0 + tid);
        };
        fetch(0);
#pragma unroll 1
        for (int p0 = 0; p0 < P; p0 += TP) {
            const int m = min(TP, P - p0);
            MBX_SYNC();                                     // previous tile fully consumed
            if (tid < m * HD) tile[tid] = r0;
            if (tid + NT < m * HD) tile[tid + NT] = r1;
            if (tid < m) tl[tid] = rl;
            MBX_SYNC();
            if (p0 + TP < P) fetch(p0 + TP);                     // in flight while this tile is consumed
            if (act) {
                // peaks of this tile with p = g (mod GP)
                int q = (g - p0 % GP + GP) % GP;
#pragma unroll 1
                for (; q < m; q += GP) visit(tile + q * HD, tl[q], p0 + q);
            }
        }
        MBX_SYNC();                                         // the tile may be reused by the caller
    }
    PROF_MARK(8);
    // merge the GP partial scans of a candidate (GP is a power of two <= 32, groups are warp-aligned)
    for (int o = GP >> 1; o > 0; o >>= 1) {
        const float os1 = __shfl_xor_sync(0xffffffffu, sc.s1, o), olo = __shfl_xor_sync(0xffffffffu, sc.lo1, o);
        const float ohi = __shfl_xor_sync(0xffffffffu, sc.hi2, o);
        const int op = __shfl_xor_sync(0xffffffffu, sc.p1, o);
        gal_merge(sc, os1, olo, ohi, op);
    }
    double gval = 0.;
    if (act && g == 0) {
        const double* zr = B + ci * ZS;
        if (sc.lo1 > sc.hi2 && sc.p1 < P) gval = gallagher_peak64<D>(gC, gRy, gw, sc.p1, zr, 1, coef);
        else {
            PROF_ADD_ANY(10, 1);                                 // profile build: count of full float64 fallbacks
            gval = -1e300;                                       // near-tie (or non-finite input): all peaks in float64
#pragma unroll 1
            for (int p = 0; p < P; ++p) gval = fmax(gval, gallagher_peak64<D>(gC, gRy, gw, p, zr, 1, coef));
        }
    }
    // hand the value to thread ci (the caller's convention: thread i owns candidate i)
    MBX_SYNC();
    PROF_MARK(9);
    double* slot = const_cast<double*>(B) + ci * ZS + D;         // the padding column of row ci
    if (act && g == 0) *slot = gval;
    MBX_SYNC();
    return (tid < n) ? B[tid * ZS + D] : 0.;
}

template <int D, class TX>
__device__ double bbob_eval_cta(const MbxInstance& h, const double* __restrict__ C, const double* __restrict__ G,
                                const TX* __restrict__ X, const int n, double* B, const unsigned char* only_mask,
                                float4* gtile, const int gtile_bytes) {
    constexpr int ZS = BbobCta<D>::ZS;
    const int tid = MBX_TID, NT = MBX_BDIM, nD = n * D;
    ConstView<D> c(C);
    const int base = h.base;
    const double bias = h.bias, ub = h.ub;
    const bool mine = tid < n && !(only_mask && !only_mask[tid]);    // this thread owns an active candidate
    const TX* xr = X + (mine ? tid : 0) * D;
    auto act = [&](const int i) { return !(only_mask && !only_mask[i]); };
    auto pen_row = [&]() {                                           // bbob.py:85-93 for candidate tid
        double p = 0.;
#pragma unroll 1
        for (int d = 0; d < D; ++d) { const double t = fabs((double)xr[d]) - ub; if (t > 0.) p += t * t; }   // max(0, .) as a compare
        return p;
    };
    double bh = 0.;
    if (mine && h.bh_coef != 0.) { const double p = pen_row(); bh = (h.bh_coef == 1.) ? p : h.bh_coef * p; }
#define MBX_EW(...)                                                          \
    for (int e = tid; e < nD; e += NT) {                                     \
        const int i = e / D, k = e - i * D;                                  \
        if (!act(i)) continue;                                               \
        double& z = B[i * ZS + k];                                           \
        __VA_ARGS__                                                          \
    }                                                                        \
    MBX_SYNC();
    // pair terms read (B[i][k], B[i][k+1]) and overwrite B[i][k]: read -> barrier -> write, round by round
#define MBX_PAIR(...)                                                        \
    for (int e0 = 0; e0 < nD; e0 += NT) {                                    \
        const int e = e0 + tid, i = e / D, k = e - i * D;                    \
        const bool on = e < nD && k + 1 < D && act(i);                       \
        double t = 0.;                                                       \
        if (on) { const double a = B[i * ZS + k], b = B[i * ZS + k + 1]; __VA_ARGS__ } \
        MBX_SYNC();                                                     \
        if (on) B[i * ZS + k] = t;                                           \
        MBX_SYNC();                                                     \
    }
    const double* zr = B + (mine ? tid : 0) * ZS;                    // this thread's row for the per-candidate passes
    // ---------------------------------------------------------------- first linear map
    if (base == 9 || base == 19) {                                   // z = M2 x + 0.5   (bbob.py:434, :703)
        MBX_EW(z = (double)X[e];)
        matvec_cta_auto<D>(c.M2, B, n, 0.5);
    } else if (base == 24) {                                         // z = R (2 sign(s) x - mu0)   (bbob.py:885-886)
        MBX_EW(z = c.cv2[k] * (double)X[e] - h.sc[0];)
        matvec_cta_auto<D>(c.R, B, n);
    } else if (base == 21 || base == 22) {                           // y = R x
        MBX_EW(z = (double)X[e];)
        matvec_cta_auto<D>(c.R, B, n);
    } else if (base != 5 && base != 8 && base != 20) {               // z = R (x - s)   (bbob.py:6-8)
        MBX_EW(z = (double)X[e] - c.s[k];)
        matvec_cta_auto<D>(c.R, B, n);
    }
    // ---------------------------------------------------------------- per-shape tails
    switch (base) {
    case 1: {                                                        // Sphere
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) { const double z = zr[k]; f += z * z; }
        }
        return f + bias + bh;
    }
    case 2: case 10: {                                               // Ellipsoidal
        MBX_EW(z = t_osz(z);)
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) { const double z = zr[k]; f += c.cv0[k] * (z * z); }
        }
        return f + bias + bh;
    }
    case 3: case 4: case 15: {                                       // Rastrigin family
        if (base == 3) { MBX_EW(z = c.cv0[k] * t_asy(t_osz(z), 0.2, c.cv1[k]);) }
        else if (base == 4) { MBX_EW(double t = t_osz(z); if (!(k & 1) && t > 0.) t *= 10.; z = t * c.cv0[k];) }
        else {
            MBX_EW(z = t_asy(t_osz(z), 0.2, c.cv1[k]);)
            matvec_cta_auto<D>(c.M2, B, n);
        }
        double sq = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) { const double z = zr[k]; sq += z * z; }
        }
        MBX_SYNC();
        MBX_EW(z = cospi(2. * z);)
        double scs = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) scs += zr[k];
        }
        const double r = 10. * ((double)D - scs) + sq;
        if (base == 4) return r + 100. * (mine ? pen_row() : 0.) + bias;
        return r + bias;
    }
    case 5: {                                                        // Linear slope (no rotation, no libm)
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int d = 0; d < D; ++d) {
                const double x = (double)xr[d];
                double z = x;
                if (x * c.s[d] > ub * ub) z = (z > 0. ? 1. : (z < 0. ? -1. : 0.)) * ub;
                const double s = c.cv0[d];
                f += ub * fabs(s) - z * s;
            }
        }
        return f + bias;
    }
    case 6: {                                                        // Attractive sector
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) {
                double z = zr[k];
                if (z * c.s[k] > 0.) z *= 100.;
                f += z * z;
            }
            f = pow(t_osz(f), 0.9);
        }
        return f + bias;
    }
    case 7: {                                                        // Step ellipsoidal
        const double zh0 = mine ? fabs(zr[0]) : 0.;
        MBX_SYNC();
        MBX_EW(z = fabs(z) > 0.5 ? floor(0.5 + z) : floor(0.5 + 10. * z) / 10.;)
        matvec_cta_auto<D>(c.M2, B, n);
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) { const double z = zr[k]; f += c.cv0[k] * (z * z); }
        }
        return 0.1 * fmax(zh0 / 1e4, f) + bh + bias;
    }
    case 8: {                                                        // Rosenbrock (R = I, no libm)
        double f = 0.;
        if (mine) {
            double a = h.sc[0] * ((double)xr[0] - c.s[0]) + 1.;
#pragma unroll 1
            for (int k = 0; k + 1 < D; ++k) {
                const double b = h.sc[0] * ((double)xr[k + 1] - c.s[k + 1]) + 1.;
                const double u = a * a - b, w = a - 1.;
                f += 100. * (u * u) + w * w;
                a = b;
            }
        }
        return f + bias + bh;
    }
    case 9: {                                                        // Rosenbrock rotated
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k + 1 < D; ++k) {
                const double a = zr[k], b = zr[k + 1];
                const double u = a * a - b, w = a - 1.;
                f += 100. * (u * u) + w * w;
            }
        }
        return f + bias;
    }
    case 11: {                                                       // Discus
        MBX_EW(z = t_osz(z);)
        double f = 0., z0 = 0.;
        if (mine) {
            z0 = zr[0];
#pragma unroll 1
            for (int k = 1; k < D; ++k) { const double z = zr[k]; f += z * z; }
        }
        return 1e6 * (z0 * z0) + f + bias;
    }
    case 12: {                                                       // Bent cigar
        MBX_EW(z = t_asy(z, 0.5, c.cv1[k]);)
        matvec_cta_auto<D>(c.R, B, n);
        double f = 0., z0 = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 1; k < D; ++k) { const double z = zr[k]; f += 1e6 * (z * z); }
            z0 = zr[0];
        }
        return z0 * z0 + f + bias;
    }
    case 13: {                                                       // Sharp ridge
        double f = 0., z0 = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 1; k < D; ++k) { const double z = zr[k]; f += z * z; }
            z0 = zr[0];
        }
        return z0 * z0 + 100. * sqrt(f) + bias;
    }
    case 14: {                                                       // Different powers
        MBX_EW(z = exp(c.cv0[k] * log(fabs(z)));)
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) f += zr[k];
        }
        return sqrt(f) + bias + bh;
    }
    case 16: {                                                       // Weierstrass
        MBX_EW(z = t_osz(z);)
        matvec_cta_auto<D>(c.M2, B, n);
        MBX_EW(z = weierstrass_inner(z);)
        double tot = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) tot += zr[k];
        }
        const double m = tot / (double)D - h.sc[0];
        return 10. * (m * m * m) + (10. / (double)D) * (mine ? pen_row() : 0.) + bias;
    }
    case 17: case 18: {                                              // Schaffers
        MBX_EW(z = t_asy(z, 0.5, c.cv1[k]);)
        matvec_cta_auto<D>(c.M2, B, n);
        MBX_PAIR(const double s = sqrt(a * a + b * b); const double sn = sin(50. * exp(0.2 * log(s))); t = sqrt(s) * (sn * sn + 1.);)
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k + 1 < D; ++k) f += zr[k];
        }
        const double m = (1. / (double)(D - 1)) * f;
        return m * m + bh + bias;
    }
    case 19: {                                                       // Griewank-Rosenbrock
        MBX_PAIR(const double u = a * a - b, w = 1. - a; const double s = 100. * (u * u) + w * w; t = s / 4000. - cos(s);)
        double f = 0.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k + 1 < D; ++k) f += zr[k];
        }
        const double fac = h.sc[0];
        return fac + fac * f / ((double)D - 1.) + bias + bh;
    }
    case 20: {                                                       // Schwefel (no rotation)
        for (int e = tid; e < nD; e += NT) {
            const int i = e / D, d = e - i * D;
            if (!act(i)) continue;
            const double z0 = c.cv2[d] * (double)X[e];
            double z = z0;
            if (d > 0) z += 0.25 * (c.cv2[d - 1] * (double)X[e - 1] - c.cv1[d - 1]);
            z = 100. * (c.cv0[d] * (z - c.cv1[d]) + c.cv1[d]);
            B[i * ZS + d] = z * sin(sqrt(fabs(z)));
        }
        MBX_SYNC();
        double f = 0., pen = 0.;
        if (mine) {
            double prev = 0.;
#pragma unroll 1
            for (int d = 0; d < D; ++d) {
                const double z0 = c.cv2[d] * (double)xr[d];
                double z = z0;
                if (d > 0) z += 0.25 * (prev - c.cv1[d - 1]);
                prev = z0;
                z = 100. * (c.cv0[d] * (z - c.cv1[d]) + c.cv1[d]);
                f += zr[d];
                const double t = fabs(z / 100.) - ub;
                if (t > 0.) pen += t * t;
            }
        }
        return 4.189828872724339 - 0.01 * (f / (double)D) + 100. * pen + bias;
    }
    case 21: case 22: {                                              // Gallagher
        const double g = gallagher_cta<D>(G, h.n_peaks, B, n, only_mask, gtile, gtile_bytes);
        if (!mine) return 0.;
        const double t = t_osz(10. - g);
        return t * t + bias + bh;
    }
    case 23: {                                                       // Katsuura
        MBX_EW(z = katsuura_term(z, k, h.sc[0]);)
        double prod = 1.;
        if (mine) {
#pragma unroll 1
            for (int k = 0; k < D; ++k) prod *= zr[k];
        }
        return prod * h.sc[1] - h.sc[1] + (mine ? pen_row() : 0.) + bias;
    }
    case 24: {                                                       // Lunacek
        MBX_EW(z = cospi(2. * z);)
        double s1 = 0., s2 = 0., scs = 0.;
        if (mine) {
#pragma unroll 1
            for (int d = 0; d < D; ++d) {
                const double xh = c.cv2[d] * (double)xr[d];
                const double a = xh - h.sc[0], b = xh - h.sc[2];
                s1 += a * a; s2 += b * b;
            }
#pragma unroll 1
            for (int k = 0; k < D; ++k) scs += zr[k];
        }
        return fmin(s1, (double)D + h.sc[1] * s2) + 10. * ((double)D - scs) + 1e4 * (mine ? pen_row() : 0.) + bias;
    }
    }
#undef MBX_EW
#undef MBX_PAIR
    return 0.;
}
