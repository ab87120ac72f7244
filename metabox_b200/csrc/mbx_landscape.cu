// mbx_landscape.cu -- landscape features of sampled walks, one CTA per lane: fitness-distance correlation, ruggedness
// (information entropy over 9 sensitivity levels), lag-1 autocorrelation, number-of-optima proxy (DEDQN backbone) and the
// dynamic FDC / ruggedness-entropy pair of the RL_HPSDE backbone.  Pure reductions over [S, D] samples + [S] costs in
// float64 (block-wide shuffle reductions, shared-memory histograms).
// Reference: optimizer/dedqn_optimizer.py:8-12 (cal_fdc), :15-51 (cal_rie), :54-61 (cal_acf), :64-76 (cal_nop);
// optimizer/rl_hpsde_optimizer.py:169-181 (DFDC), :184-216 (DRIE).  numpy twin: oracle/landscape.py.
#include "mbx_common.cuh"
#include "mbx_launch.h"

#define LS_NT 128
#define LS_MAX_S 512

__device__ __forceinline__ double ls_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < LS_NT / 32; ++k) s += red[k];
    return s;
}
__device__ __forceinline__ double ls_block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < LS_NT / 32; ++k) s = fmax(s, red[k]);
    return s;
}
// entropy -sum p log p / log 6 of six counts; `zero_as`: what an empty class is replaced by before the division by `den`
__device__ __forceinline__ double ls_entropy(const int* cnt, const double den, const double zero_count, const double pmin) {
    double e = 0.;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double p = (cnt[k] == 0 ? zero_count : (double)cnt[k]) / den;
        if (p < pmin) p = pmin;
        e += p * log(p) / log(6.);
    }
    return -e;
}

__global__ void __launch_bounds__(LS_NT)
landscape_kernel(const double* __restrict__ sample, const double* __restrict__ cost, const double* __restrict__ ref, const int S,
                 const int D, double* __restrict__ out) {
    __shared__ double sc[LS_MAX_S], sd[LS_MAX_S], sr[LS_MAX_S], sfs[LS_MAX_S], red[8];
    __shared__ int hist[2][9][6], sbest;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* X = sample + (size_t)b * S * D;
    const double* R = ref ? ref + (size_t)b * D : nullptr;
    for (int i = tid; i < S; i += LS_NT) sc[i] = cost[(size_t)b * S + i];
    for (int k = tid; k < 2 * 9 * 6; k += LS_NT) (&hist[0][0][0])[k] = 0;
    __syncthreads();
    // ---- best = argmin(cost), first occurrence -------------------------------------------------------------------------------
    if (tid == 0) {
        int bi = 0;
        for (int i = 1; i < S; ++i) if (sc[i] < sc[bi]) bi = i;
        sbest = bi;
    }
    __syncthreads();
    const int best = sbest;
    // ---- distances to the best sample (sd) and to the reference point (sr) -------------------------------------------------------
    for (int i = tid; i < S; i += LS_NT) {
        double a = 0., r = 0.;
        for (int d = 0; d < D; ++d) {
            const double x = X[(size_t)i * D + d];
            const double u = x - X[(size_t)best * D + d], v = x - (R ? R[d] : X[(size_t)best * D + d]);
            a += u * u; r += v * v;
        }
        sd[i] = sqrt(a); sr[i] = sqrt(r);
    }
    __syncthreads();
    // ---- cal_fdc / cal_acf (dedqn_optimizer.py:8-12, :54-61) -------------------------------------------------------------------
    double pf = 0., pd = 0.;
    for (int i = tid; i < S; i += LS_NT) { pf += sc[i]; pd += sd[i]; }
    const double mf = ls_block_sum(pf, red) / S, md = ls_block_sum(pd, red) / S;
    double cfd = 0., vf = 0., vd = 0., ac = 0.;
    for (int i = tid; i < S; i += LS_NT) {
        const double a = sc[i] - mf, c = sd[i] - md;
        cfd += a * c; vf += a * a; vd += c * c;
        if (i + 1 < S) ac += a * (sc[i + 1] - mf);
    }
    cfd = ls_block_sum(cfd, red); vf = ls_block_sum(vf, red); vd = ls_block_sum(vd, red); ac = ls_block_sum(ac, red);
    const double fdc = (cfd / S) / ((vd / S) * (vf / S) + 1e-6);
    const double acf = ac / (vf + 1e-6);
    // ---- cal_nop (:64-76): costs ordered by distance (stable), count of descents -----------------------------------------------
    for (int i = tid; i < S; i += LS_NT) {
        const double di = sd[i];
        int r = 0;
        for (int j = 0; j < S; ++j) { const double dj = sd[j]; r += (dj < di || (dj == di && j < i)) ? 1 : 0; }
        sfs[r] = sc[i];
    }
    __syncthreads();
    double nd = 0.;
    for (int i = tid; i + 1 < S; i += LS_NT) nd += (sfs[i + 1] < sfs[i]) ? 1. : 0.;
    const double nop = ls_block_sum(nd, red) / S;
    // ---- cal_rie (:15-51) and DRIE (rl_hpsde_optimizer.py:184-216): symbol-pair histograms at 9 sensitivity levels ----------------
    double mx = 0., mxa = 0.;
    for (int i = tid; i + 1 < S; i += LS_NT) { const double d = sc[i + 1] - sc[i]; mx = fmax(mx, d); mxa = fmax(mxa, fabs(d)); }
    const double eps_star = ls_block_max(mx, red), e_star = ls_block_max(mxa, red);
    for (int i = tid; i + 2 < S; i += LS_NT) {
        const double d0 = sc[i + 1] - sc[i], d1 = sc[i + 2] - sc[i + 1];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            // DEDQN levels: eps* / 2^k (k < 8), 0 ; classes (-1,0) (-1,1) (0,1) (0,-1) (1,-1) else
            const double eps = k < 8 ? eps_star / (double)(1 << k) : 0.;
            const int a = d0 < -eps ? -1 : (d0 > eps ? 1 : 0), c = d1 < -eps ? -1 : (d1 > eps ? 1 : 0);
            int cls = 5;
            if (a == -1 && c == 0) cls = 0; else if (a == -1 && c == 1) cls = 1; else if (a == 0 && c == 1) cls = 2;
            else if (a == 0 && c == -1) cls = 3; else if (a == 1 && c == -1) cls = 4;
            atomicAdd(&hist[0][k][cls], 1);
            // RL_HPSDE levels: scale * e*, scale in {0, 1/128 ... 1}; classes (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0), no "else"
            const double sc2 = k == 0 ? 0. : e_star / (double)(1 << (8 - k));
            const int a2 = d0 < -sc2 ? -1 : (sc2 < d0 ? 1 : 0), c2 = d1 < -sc2 ? -1 : (sc2 < d1 ? 1 : 0);
            int cl2 = -1;
            if (a2 == -1 && c2 == 0) cl2 = 0; else if (a2 == -1 && c2 == 1) cl2 = 1; else if (a2 == 0 && c2 == -1) cl2 = 2;
            else if (a2 == 0 && c2 == 1) cl2 = 3; else if (a2 == 1 && c2 == -1) cl2 = 4; else if (a2 == 1 && c2 == 0) cl2 = 5;
            if (cl2 >= 0) atomicAdd(&hist[1][k][cl2], 1);
        }
    }
    __syncthreads();
    double e1 = -1e300, e2 = -1e300;
    if (tid < 9) {
        e1 = ls_entropy(hist[0][tid], (double)S, (double)S, 0.);                 // freq[freq == 0] = len ; freq /= len
        e2 = ls_entropy(hist[1][tid], (double)(S - 1), 0., 1e-15);               // prob /= rw_steps ; prob[prob < 1e-15] = 1e-15
    }
    const double rie = ls_block_max(e1, red), drie = ls_block_max(e2, red);
    // ---- DFDC (rl_hpsde_optimizer.py:169-174): Pearson r of cost vs distance to ref over the walk without its first point ----------
    double qc = 0., qd = 0.;
    for (int i = 1 + tid; i < S; i += LS_NT) { qc += sc[i]; qd += sr[i]; }
    const double mc2 = ls_block_sum(qc, red) / (S - 1), md2 = ls_block_sum(qd, red) / (S - 1);
    double cv = 0., vc = 0., vr = 0.;
    for (int i = 1 + tid; i < S; i += LS_NT) { const double a = sc[i] - mc2, c = sr[i] - md2; cv += a * c; vc += a * a; vr += c * c; }
    cv = ls_block_sum(cv, red); vc = ls_block_sum(vc, red); vr = ls_block_sum(vr, red);
    const double dfdc = (cv / (S - 1)) / (sqrt(vc / (S - 1)) * sqrt(vr / (S - 1)));
    if (tid == 0) {
        double* o = out + (size_t)b * 6;
        o[0] = fdc; o[1] = rie; o[2] = acf; o[3] = nop; o[4] = dfdc; o[5] = drie;
    }
}

int mbx_launch_landscape(const double* sample, const double* cost, const double* ref, int B, int S, int D, double* out,
                         cudaStream_t st) {
    landscape_kernel<<<B, LS_NT, 0, st>>>(sample, cost, ref, S, D, out);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
