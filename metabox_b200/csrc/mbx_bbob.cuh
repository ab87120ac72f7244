// mbx_bbob.cuh -- device building blocks of the fitness functions: instance-constant views, the BBOB transforms
// (T_osz, T_asy, penalty), the Gallagher float32 peak screening state + exact float64 peak value, the noise models, and the
// protein-docking energy (one warp per candidate).  float64 arithmetic in the reference's operation order
// (reference: problem/bbob.py:6-8,51-93,108-146,149-890; problem/protein_docking.py:28-48).
// The 24 BBOB shapes themselves are evaluated CTA-cooperatively in mbx_bbob_cta.cuh.
#pragma once
#include "mbx_common.cuh"

struct XF32 {
    const float* p;
    __device__ __forceinline__ double operator()(int d) const { return (double)p[d]; }
};
struct XF64 {
    const double* p;
    __device__ __forceinline__ double operator()(int d) const { return p[d]; }
};

template <int D>
struct ConstView {          // shared-memory views of one BBOB instance
    const double *s, *R, *M2, *cv0, *cv1, *cv2;
    __device__ __forceinline__ ConstView(const double* C)
        : s(C), R(C + D), M2(C + D + D * D), cv0(C + D + 2 * D * D), cv1(C + 2 * D + 2 * D * D), cv2(C + 3 * D + 2 * D * D) {}
    static constexpr int kDoubles = 4 * D + 2 * D * D;
};

__device__ __forceinline__ double t_osz(double x) {       // bbob.py:51-67
    double ax = fabs(x);
    double t = log(ax) * 10.;                              // reference: / 0.1 (<= 1 ulp apart)
    double c1 = x > 0. ? 1.0 : 0.55, c2 = x > 0. ? 0.79 : 0.31;
    // reference: (exp(t + 0.49(sin c1 t + sin c2 t)))^0.1 ; exp(0.1 u) == exp(u)^0.1 up to a few ulp and saves the pow
    double r = exp((t + 0.49 * (sin(c1 * t) + sin(c2 * t))) * 0.1);
    return x > 0. ? r : (x < 0. ? -r : x);
}
__device__ __forceinline__ double t_asy(double x, double beta, double frac) {   // bbob.py:70-82
    return x > 0. ? exp((1. + beta * frac * sqrt(x)) * log(x)) : x;      // x^e as exp(e ln x): a few ulp, far inside 1e-6
}
template <int D, class XA>
__device__ __forceinline__ double pen_x(const XA& X, double ub) {               // bbob.py:85-93
    double p = 0.;
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const double t = fabs(X(d)) - ub;                  // max(0, .) as a compare: 0 * 0 added to p changes nothing
        if (t > 0.) p += t * t;
    }
    return p;
}
// Gallagher: max_p w_p exp(-(1/2D) sum_k C_pk (y_k - (R y_p)_k)^2), y = R x in zs (bbob.py:795-799).
// exp is monotone, so the maximum is attained at argmax_p s_p, s_p = ln w_p - acc_p / 2D.  The P x D scan runs in float32
// (packed f32x2, constants pre-rounded by the host).  Error budget of the float32 acc_p: each z_k = y_k - c_k is off by at
// most E = 1.3e-7 (max|y| + 5 sqrt(D) + 1) (input roundings + the subtraction; |c| <= |y_p|_2 <= 5 sqrt(D)), which moves
// sum C z^2 by <= 2 E sqrt(sum C) sqrt(acc) + E^2 sum C (Cauchy-Schwarz; C_pk <= sqrt(1000) so sum C <= 31.7 D); products and
// the non-negative summation add <= (D + 4) 2^-24 acc.  delta_p = |coef| (K1 sqrt(acc) + 4e-6 acc) + 3e-6 bounds the error
// of s_p with margin (ln w, coef and the final fma roundings are ~1e-7 |s|).  The scan keeps the best
// s_p and the highest upper bound among the others; when the best lower bound clears that upper bound, the float64 value
// is computed for that peak alone with the reference's operation order -- bit-identical to evaluating all P peaks in
// float64.  Otherwise (near-tie) the candidate falls back to the full float64 loop.
template <int D>
__device__ __forceinline__ double gallagher_peak64(const double* __restrict__ gC, const double* __restrict__ gRy, const double* __restrict__ gw,
                                                   const int p, const double* zs, const int zst, const double coef) {
    double acc = 0.;
#pragma unroll 2
    for (int k = 0; k < D; ++k) {
        const double z = zs[k * zst] - __ldg(gRy + p * D + k);
        acc += __ldg(gC + p * D + k) * (z * z);
    }
    return __ldg(gw + p) * exp(coef * acc);
}
struct GalScan {            // running (best s, its lower bound, its peak) and the highest upper bound among the others
    float s1 = -INFINITY, lo1 = -INFINITY, hi2 = -INFINITY;
    int p1 = 0;
    __device__ __forceinline__ void visit(const float acc, const float lnw, const float coef32, const float K1, const int p) {
        const float s = fmaf(coef32, acc, lnw);
        const float dl = fmaf(K1, sqrtf(acc), fmaf(-4e-6f * coef32, acc, 3e-6f));
        const float hi = s + dl, lo = s - dl;
        if (s > s1) { hi2 = fmaxf(hi2, s1 + (s1 - lo1)); s1 = s; lo1 = lo; p1 = p; }
        else hi2 = fmaxf(hi2, hi);
    }
};
// Noise models bbob.py:108-146.  u1,u2 uniforms / n1,n2 normals in the reference's draw order.
__device__ __forceinline__ double bbob_noise(const MbxInstance& h, double ftrue, double d0, double d1, double d2) {
    if (h.noise == MBX_NOISE_NONE) return ftrue;
    double u = ftrue - h.optimum, v;
    if (h.noise == MBX_NOISE_GAUSS) {
        v = u * exp(h.n_beta * d0);
    } else if (h.noise == MBX_NOISE_UNIFORM) {
        double a = (h.n_beta == 1.) ? d0 : pow(d0, h.n_beta);
        v = u * a * fmax(1., pow(1e9 / (u + 1e-99), h.n_alpha * (0.49 + 1. / (double)h.dim) * d1));
    } else {
        v = u + h.n_alpha * fmax(0., 1e3 + (d0 < h.n_p ? 1. : 0.) * d1 / (fabs(d2) + 1e-199));
    }
    return (u >= 1e-8) ? (v + h.optimum + 1.01 * 1e-8) : ftrue;
}
// Philox draws for one candidate of a noisy evaluation (numpy twin: oracle/rlepso.py philox_noise)
__device__ __forceinline__ void noise_draws(const Philox& ph, int kind, uint32_t i, uint32_t gen, uint32_t lane, uint32_t tag,
                                            double& d0, double& d1, double& d2) {
    d0 = d1 = d2 = 0.;
    if (kind == MBX_NOISE_NONE) return;
    uint4 w = ph(i, gen, lane, tag);
    if (kind == MBX_NOISE_GAUSS) { double t; box_muller(w.x, w.y, d0, t); }
    else if (kind == MBX_NOISE_UNIFORM) { d0 = u01(w.x); d1 = u01(w.y); }
    else { d0 = u01(w.x); box_muller(w.y, w.z, d1, d2); }
}

// ------------------------------------------------------------------------------------- protein
// Protein_Docking.func protein_docking.py:28-48.  Constants in smem: coor_init[300] | basis[12*300] |
// inv_sqrt_eig[12] | q[100] | e4[100] (= e_i^(1/4), so sqrt(e_ij) = e4_i*e4_j) | r[100].
// ONE WARP per candidate: the 300 coordinates are built cooperatively into a per-warp smem block
// (SoA cx|cy|cz|p2, 400 doubles), then the 4950 unordered atom pairs are strided over the 32 lanes
// (rows i and 99-i are paired so every lane-step is useful) and the partial sums are combined with a
// shuffle reduction.  The reference sums all ordered pairs; the matrix is symmetric and the diagonal
// is masked out (d = 0.1 < 0.11), so energy = 2 * sum_{i<j} / 100.
#define MBX_PROT_ATOMS 100
#define MBX_PROT_DOUBLES (300 + 12 * 300 + 12 + 300)
#define MBX_PROT_WARP_SCRATCH 400
// Static pair pruning: the host lists the atom pairs (i < j) that can come within the 9 A cut-off for SOME x in
// [lb, ub]^12 (initial distance minus both atoms' maximal displacement <= 9, problem.py Protein_Docking.const_block); the
// other pairs (46-64 % for the shipped complexes) are masked out by the reference for every in-bounds candidate, so
// skipping them is exact.  Candidates outside the bounds (only reachable through mbx_eval) take all 4950 pairs.
// sqrt and the two divisions by d are replaced by one rsqrt: d = d2 * rsqrt(d2), r/d = r * rsqrt(d2) (~1 ulp apart).
template <class XA>
__device__ double protein_energy_warp(const double* __restrict__ C, const XA& X, double* ws, int lane,
                                      const unsigned short* __restrict__ plist = nullptr, const int n_pairs = 0,
                                      const double bound = 1.5) {
    const double* coor0 = C;
    const double* basis = C + 300;
    const double* ise = C + 300 + 3600;
    const double* q = ise + 12;
    const double* e4 = q + 100;
    const double* r = e4 + 100;
    double xs[12];
    bool inb = true;
#pragma unroll
    for (int k = 0; k < 12; ++k) { const double xk = X(k); inb = inb && (fabs(xk) <= bound); xs[k] = xk * ise[k]; }
    __syncwarp();
    for (int a = lane; a < 300; a += 32) {
        double acc = 0.;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc = fma(xs[k], basis[k * 300 + a], acc);
        int atom = a / 3, comp = a - atom * 3;
        ws[comp * 100 + atom] = acc + coor0[a];
    }
    __syncwarp();
    for (int a = lane; a < 100; a += 32) {
        double x = ws[a], y = ws[100 + a], z = ws[200 + a];
        ws[300 + a] = x * x + y * y + z * z;
    }
    __syncwarp();
    auto pair_term = [&](const int i, const int j) -> double {
        const double p3 = fma(ws[i], ws[j], fma(ws[100 + i], ws[100 + j], ws[200 + i] * ws[200 + j]));
        const double d2 = ws[300 + i] - 2. * p3 + ws[300 + j] + 0.01;
        if (!(d2 > 0.0120 && d2 < 81.001)) return 0.;             // certainly outside (0.11, 9): no sqrt needed
        const double inv = rsqrt(d2), d = d2 * inv;
        const bool near = (d > 0.11) && (d < 7.0);
        const bool far = (d > 7.0) && (d < 9.0);
        if (!(near || far)) return 0.;
        const double rr = (r[i] + r[j]) / 2.;
        const double t1 = rr * inv;
        const double t2 = t1 * t1, t6 = t2 * t2 * t2;
        const double coeff = (q[i] * q[j]) * (0.25 * inv) + (e4[i] * e4[j]) * (t6 * t6 - t6);
        double term = 10. * coeff;
        if (far) { const double g = 9. - d; term = term * (g * g * (-12. + 2. * d) / 8.); }
        return term;
    };
    double part = 0.;
    if (inb && plist != nullptr && n_pairs > 0) {
#pragma unroll 1
        for (int p = lane; p < n_pairs; p += 32) {
            const unsigned int ij = __ldg(plist + p);
            part += pair_term((int)(ij >> 8), (int)(ij & 255u));
        }
    } else {
#pragma unroll 1
        for (int p = lane; p < 4950; p += 32) {
            int cidx = p / 99, t = p - cidx * 99;
            int i, j;
            if (t < 99 - cidx) { i = cidx; j = cidx + 1 + t; }
            else { i = 99 - cidx; j = i + 1 + (t - (99 - cidx)); }
            part += pair_term(i, j);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncwarp();
    return 2. * part / (double)MBX_PROT_ATOMS;
}
