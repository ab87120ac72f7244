// mbx_deddqn.cu -- fused DE-DDQN step (DE family, ONE individual per step): generation bookkeeping ->
// mutation by the chosen operator (rand/1, rand/2, rand-to-best/2, cur-to-rand/1; F = 0.5, Cr = 1) ->
// clip -> fitness of the single trial -> reward -> operator-credit history (N_tot / N_succ / OM per
// generation slot, window OM_W) -> greedy `<=` replacement, gbest / gworst / stagnation -> cost-curve
// log -> the 99 state features of the next individual.  One CTA per lane; the single trial is evaluated by the
// whole CTA (one dimension per thread, mbx_bbob_cta.cuh).
// Reference: optimizer/de_ddqn_optimizer.py:46-74 (init_population), :76-129 (__get_state), :131-220
// (update); operators/mutate.py:36-39,53-56,124-127,178-181; crossover.py:6-18; boundary_control.py:5-9.
// numpy twin: oracle/deddqn.py (documents the reference quirks that are reproduced here).
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#ifdef MBX_PROFILE
extern "C" int mbx_prof_read_deddqn(unsigned long long* h16, int reset) {
    if (cudaMemcpyFromSymbol(h16, g_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return MBX_ECUDA;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return MBX_ECUDA;
    }
    return MBX_OK;
}
#endif

#define DD_GEN_MAX 10
#define DD_W 50
#define DD_WS 8           // doubles per window entry: action, om0..om3, trial_cost, seq, pad
#define DD_NFEAT 99
// ints[] slots
#define DD_I_POINTER 0
#define DD_I_GEN 1
#define DD_I_STAG 2
#define DD_I_PROW 3       // >= 0: X_prebest is still a view of this row of X
#define DD_I_REBOUND 4    // X_gbest has been re-bound to a trial
#define DD_I_WLEN 5
#define DD_I_WSEQ 6       // next window sequence number
#define DD_I_ERR 7        // 1: an action outside 0..3 was seen (reference raises ValueError('Action error'))

template <int D, int KIND>
struct DdqnSmem {
    static constexpr int N = MBX_MAX_NP;
    static constexpr int kConst = EvalTraits<D, KIND>::kConstDoubles;
    static constexpr int kScratch = (KIND == MBX_KIND_PROTEIN) ? MBX_PROT_WARP_SCRATCH : D + 1;  // one candidate
    static constexpr size_t oHdr = 0;
    static constexpr size_t oConst = sizeof(MbxInstance);
    static constexpr size_t oScr = oConst + sizeof(double) * kConst;
    static constexpr size_t oWin = oScr + sizeof(double) * kScratch;          // window copy [W*WS] f64
    static constexpr size_t oRedD = oWin + sizeof(double) * DD_W * DD_WS;     // f64 reduction scratch [16]
    static constexpr size_t oCost = oRedD + sizeof(double) * 16;              // costs f32 [N]
    static constexpr size_t oTrial = oCost + sizeof(float) * N;               // trial f32 [MAX_DIM]
    static constexpr size_t oMisc = oTrial + sizeof(float) * MBX_MAX_DIM;     // misc floats / ints [32]
    static constexpr size_t kBytes = ((oMisc + 128 + 15) / 16) * 16;
};

__device__ __forceinline__ double dd_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < MBX_NT / 32; ++k) s += red[k];
    return s;
}
// stable rank of cost[i] among cost[0..NP) (NaN last)
__device__ __forceinline__ int dd_rank(const float* c, int i, int NP) {
    const float ci = c[i];
    int r = 0;
    if (isnan(ci)) {
        for (int j = 0; j < NP; ++j) { const float cj = c[j]; r += (!isnan(cj) || j < i) ? 1 : 0; }
        return r;
    }
    int lt = 0, eq = 0;                                   // a NaN c[j] fails both comparisons: it sorts after ci
#pragma unroll 4
    for (int j = 0; j < i; ++j) { const float cj = c[j]; lt += cj < ci ? 1 : 0; eq += cj == ci ? 1 : 0; }
#pragma unroll 4
    for (int j = i + 1; j < NP; ++j) lt += c[j] < ci ? 1 : 0;
    return lt + eq;
}

// ---- __get_state (de_ddqn_optimizer.py:76-129): all threads call; needs sc[] = current costs in smem ------------------
template <int D>
__device__ void dd_features(const MbxRunConfig& cfg, const MbxDdqnState& S, int b, const float* sc, double* red, double* swin,
                            int p, int gen, int stag, int prow, int wlen, int fes, float cgb, float cgw, float cpb,
                            const int* r, double lb, double ub, float* feat) {
    const int NP = cfg.NP, tid = threadIdx.x;
    const double gb = (double)cgb, span = (double)cgw - (double)cgb;
    // mean / std of the costs (np.mean, np.std)
    const double ci = tid < NP ? (double)sc[tid] : 0.;
    const double mean = dd_block_sum(ci, red) / (double)NP;
    const double dv = tid < NP ? (ci - mean) : 0.;
    const double var = dd_block_sum(dv * dv, red) / (double)NP;
    const float* X = S.X + (size_t)b * NP * D;
    const double cp = (double)sc[p];
    const double max_dist = sqrt((double)D * ((ub - lb) * (ub - lb)));
    if (tid == 0) {
        feat[0] = (float)((cp - gb) / span);
        feat[1] = (float)((mean - gb) / span);
        feat[2] = (float)(sqrt(var) / (span / 2.));
        feat[3] = (float)((double)(cfg.maxFEs - fes) / (double)cfg.maxFEs);
        feat[4] = 1.f;
        feat[5] = (float)((double)stag / (double)cfg.maxFEs);
        feat[17] = (float)((cp - (double)cpb) / span);
    }
    if (tid < 5) feat[12 + tid] = (float)((cp - (double)sc[r[tid]]) / span);
    // 7 distances: X[p] to X[r_j] (6..10), X_prebest (11), X_gbest (18); warp w handles targets w, w+4
    {
        const int w = tid >> 5, l = tid & 31;
        for (int t = w; t < 7; t += MBX_NT / 32) {
            const float* Y = t < 5 ? X + (size_t)r[t] * D
                                   : (t == 5 ? (prow >= 0 ? X + (size_t)prow * D : S.X_prebest + (size_t)b * D)
                                             : S.X_gbest + (size_t)b * D);
            double acc = 0.;
            for (int d = l; d < D; d += 32) { const double df = (double)X[(size_t)p * D + d] - (double)Y[d]; acc += df * df; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (l == 0) feat[t < 5 ? 6 + t : (t == 5 ? 11 : 18)] = (float)(sqrt(acc) / max_dist);
        }
    }
    // operator-credit aggregates: thread (op,m) = tid < 16
    if (tid < 16) {
        const int op = tid >> 2, m = tid & 3;
        const int G = min(DD_GEN_MAX, gen);
        const int32_t* nt = S.n_tot + ((size_t)b * 4 + op) * DD_GEN_MAX;
        const int32_t* ns = S.n_succ + (((size_t)b * 4 + op) * 4 + m) * DD_GEN_MAX;
        const double* os = S.om_sum + (((size_t)b * 4 + op) * 4 + m) * DD_GEN_MAX;
        const double* ox = S.om_max + (((size_t)b * 4 + op) * 4 + m) * DD_GEN_MAX;
        double f19 = 0., f35 = 0., f67 = 0.;
        int tot = 0;
#pragma unroll
        for (int g0 = 0; g0 < DD_GEN_MAX; g0 += 5) {                      // loads of 5 slots in flight together
            int nv[5], kv[5]; double ov[5], xv[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const bool on = g0 + j < G;
                nv[j] = on ? nt[g0 + j] : 0; kv[j] = on ? ns[g0 + j] : 0; ov[j] = on ? os[g0 + j] : 0.; xv[j] = on ? ox[g0 + j] : 0.;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (g0 + j < G) {
                    const int n = nv[j], k = kv[j];
                    tot += n;
                    if (n > 0) f19 += (double)k / (double)n;
                    if (k > 0) { f35 += ov[j]; f67 += xv[j]; }
                }
            }
        }
        if (tot > 0) f35 = f35 / (double)tot;
        double f51 = 0.;
        if (gen >= 2) {
            const int dn = nt[0] - nt[1];
            if (dn != 0 && ns[0] > 0 && ns[1] > 0) f51 = (ox[0] - ox[1]) / (ox[1] * fabs((double)dn));
        }
        // window sums (features 83..98): entries whose action == op, in list (sequence) order
        double f83 = 0.;
        {
            // entries are unordered slots (sequence number in [6]); the sum order differs from the list order by rounding only
            for (int w = 0; w < wlen; ++w)
                if ((int)swin[w * DD_WS] == op) f83 += swin[w * DD_WS + 1 + m];
        }
        feat[19 + tid] = (float)f19; feat[35 + tid] = (float)f35; feat[51 + tid] = (float)f51;
        feat[67 + tid] = (float)f67; feat[83 + tid] = (float)f83;
    }
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
deddqn_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   MbxRunConfig cfg, MbxDdqnState S, MbxDdqnTape T, float* __restrict__ feature) {
    using L = PopSmem<D, KIND>;                                            // NP candidates at once: the population layout
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw + L::oHdr);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oA);
    double* red = reinterpret_cast<double*>(smem_raw + L::oPer);
    double* swin = red + 16;
    float* xs = reinterpret_cast<float*>(smem_raw + L::oXs);
    float* sc = reinterpret_cast<float*>(smem_raw + L::oF);
    float* redv = reinterpret_cast<float*>(smem_raw + L::oRed);
    int* redi = reinterpret_cast<int*>(smem_raw + L::oRed + 32);
    __shared__ int sr[5];
    const int b = blockIdx.x, NP = cfg.NP, tid = threadIdx.x;
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, sh, sC);
    __syncthreads();
    const MbxInstance& h = *sh;
    Philox ph(cfg.seed);
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    const double lb = h.lb, ub = h.ub;
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    for (int e = tid; e < NP * D; e += MBX_NT) {
        const double u = cfg.use_tape ? T.init_u[base + e] : u01(ph((uint32_t)e, MBX_GEN_INIT, lane, MBX_TAG_INIT).x);
        const float p = (float)(u * (ub - lb) + lb);                       // :48
        xs[e] = p; S.X[base + e] = p;
    }
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, sC, consts + gi->off, xs, NP, 0, scratch, sc, true, ph, MBX_GEN_INIT, lane,
                                         MBX_TAG_NOISE, cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP);
    __syncthreads();
    float v = INFINITY; int ix = tid;
    if (tid < NP) { v = sc[tid]; S.cost[pb + tid] = v; }
    block_argmin<MBX_NT>(v, ix, redv, redi);
    const float worst = block_max<MBX_NT>(tid < NP ? sc[tid] : -INFINITY, redv);
    for (int d = tid; d < D; d += MBX_NT) { S.X_gbest[(size_t)b * D + d] = xs[ix * D + d]; S.X_prebest[(size_t)b * D + d] = xs[ix * D + d]; }
    // zero the history
    for (int k = tid; k < 4 * DD_GEN_MAX; k += MBX_NT) S.n_tot[(size_t)b * 4 * DD_GEN_MAX + k] = 0;
    for (int k = tid; k < 16 * DD_GEN_MAX; k += MBX_NT) {
        S.n_succ[(size_t)b * 16 * DD_GEN_MAX + k] = 0; S.om_sum[(size_t)b * 16 * DD_GEN_MAX + k] = 0.; S.om_max[(size_t)b * 16 * DD_GEN_MAX + k] = 0.;
    }
    for (int k = tid; k < DD_W * DD_WS; k += MBX_NT) S.win[(size_t)b * DD_W * DD_WS + k] = 0.;
    if (tid < 5) {
        int rj;
        if (cfg.use_tape) rj = T.r5[(size_t)b * 5 + tid];
        else { uint4 w = ph((uint32_t)(tid >> 2), MBX_GEN_INIT, lane, MBX_TAG_DDQN); rj = bounded(tid == 0 || tid == 4 ? w.x : (tid == 1 ? w.y : (tid == 2 ? w.z : w.w)), NP); }
        sr[tid] = rj; S.r[(size_t)b * 5 + tid] = rj;
    }
    int32_t* ints = S.ints + (size_t)b * 8;
    if (tid == 0) {
        ints[DD_I_POINTER] = 0; ints[DD_I_GEN] = 0; ints[DD_I_STAG] = 0; ints[DD_I_PROW] = ix; ints[DD_I_REBOUND] = 0;
        ints[DD_I_WLEN] = 0; ints[DD_I_WSEQ] = 0; ints[DD_I_ERR] = 0;
        S.c3[(size_t)b * 3] = v; S.c3[(size_t)b * 3 + 1] = worst; S.c3[(size_t)b * 3 + 2] = v;
        S.fes[b] = NP; S.log_index[b] = 1; S.curve[(size_t)b * MBX_CURVE_LEN] = v; S.curve_len[b] = 1; S.done[b] = 0;
    }
    __syncthreads();
    dd_features<D>(cfg, S, b, sc, red, swin, 0, 0, 0, ix, 0, NP, v, worst, v, sr, lb, ub, feature + (size_t)b * DD_NFEAT);
}

// =====================================================================================================
template <int D, int KIND>
#ifndef MBX_DEDDQN_MIN_CTAS
#define MBX_DEDDQN_MIN_CTAS 8        // compile-time knob: minimum CTAs per SM of the step kernel (A/B builds: python -m metabox_b200.build --variant)
#endif
__global__ void __launch_bounds__(MBX_NT, MBX_DEDDQN_MIN_CTAS)
deddqn_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   MbxRunConfig cfg, MbxDdqnState S, MbxDdqnTape T, const int32_t* __restrict__ action,
                   float* __restrict__ feature, float* __restrict__ reward, float* __restrict__ dbg) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }      // finished lane: state frozen, reward 0
    PROF_INIT
    using L = DdqnSmem<D, KIND>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw + L::oHdr);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oScr);
    double* swin = reinterpret_cast<double*>(smem_raw + L::oWin);
    double* red = reinterpret_cast<double*>(smem_raw + L::oRedD);
    float* sc = reinterpret_cast<float*>(smem_raw + L::oCost);
    float* trial = reinterpret_cast<float*>(smem_raw + L::oTrial);
    float* mf = reinterpret_cast<float*>(smem_raw + L::oMisc);             // [0] trial cost, [1],[2] median pair
    int* mi = reinterpret_cast<int*>(smem_raw + L::oMisc + 32);            // [0..4] r (old), [5..9] r (new), [10] replaced, [11] new gbest
    const int NP = cfg.NP, tid = threadIdx.x;
    const int a = action[b];
    int32_t* ints = S.ints + (size_t)b * 8;
    if (a < 0 || a > 3) { if (tid == 0) ints[DD_I_ERR] = 1; return; }      // reference: ValueError('Action error') (:153)
    const MbxInstance* gi = inst + lane_inst[b];
    // one candidate per step touches each constant about once: BBOB constants are read straight from global memory (L1/L2
    // resident, shared by all lanes of the instance) instead of staging 15 KB per CTA per step; the protein energy re-reads
    // its basis per atom and keeps the staged copy
    const double* cC = sC;
    if (KIND == MBX_KIND_PROTEIN) stage_instance<D, KIND>(gi, consts, sh, sC);
    else {
        const int* hs = reinterpret_cast<const int*>(gi);
        int* hd = reinterpret_cast<int*>(sh);
        for (int k = tid; k < (int)(sizeof(MbxInstance) / 4); k += MBX_NT) hd[k] = hs[k];
        cC = consts + gi->off;
    }
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    float* X = S.X + base;
    float* Xg = S.X_gbest + (size_t)b * D;
    float* Xp = S.X_prebest + (size_t)b * D;
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    const int p = ints[DD_I_POINTER];
    int gen = ints[DD_I_GEN], stag = ints[DD_I_STAG], prow = ints[DD_I_PROW], rebound = ints[DD_I_REBOUND];
    int wlen = ints[DD_I_WLEN], wseq = ints[DD_I_WSEQ];
    const int fes0 = S.fes[b];
    float cgb = S.c3[(size_t)b * 3], cgw = S.c3[(size_t)b * 3 + 1];
    const float cpb = S.c3[(size_t)b * 3 + 2];                             // never refreshed (:135)
    if (tid < NP) sc[tid] = S.cost[pb + tid];
    if (tid < 5) mi[tid] = S.r[(size_t)b * 5 + tid];
    for (int k = tid; k < wlen * DD_WS; k += MBX_NT) swin[k] = S.win[(size_t)b * DD_W * DD_WS + k];
    // ---- generation boundary (:132-142) ------------------------------------------------------------------------
    if (p == 0) {
        if (rebound) {                                                     // X_prebest = X_gbest (a trial array by now)
            prow = -1;
            for (int d = tid; d < D; d += MBX_NT) Xp[d] = Xg[d];
        }
        gen += 1;
        // appendleft(0) on the maxlen-10 deques: slot g <- slot g-1
        int32_t* nt = S.n_tot + (size_t)b * 4 * DD_GEN_MAX;
        int32_t* ns = S.n_succ + (size_t)b * 16 * DD_GEN_MAX;
        double* os = S.om_sum + (size_t)b * 16 * DD_GEN_MAX;
        double* ox = S.om_max + (size_t)b * 16 * DD_GEN_MAX;
        int vt = 0, vs[2] = {0, 0}; double vo[2] = {0., 0.}, vx[2] = {0., 0.};
        if (tid < 4 * DD_GEN_MAX) vt = (tid % DD_GEN_MAX) ? nt[tid - 1] : 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = tid + q * MBX_NT;
            if (k < 16 * DD_GEN_MAX && (k % DD_GEN_MAX)) { vs[q] = ns[k - 1]; vo[q] = os[k - 1]; vx[q] = ox[k - 1]; }
        }
        __syncthreads();
        if (tid < 4 * DD_GEN_MAX) nt[tid] = vt;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = tid + q * MBX_NT;
            if (k < 16 * DD_GEN_MAX) { ns[k] = vs[q]; os[k] = vo[q]; ox[k] = vx[q]; }
        }
    }
    __syncthreads();
    PROF_MARK(0);
    const MbxInstance& h = *sh;
    const double lb = h.lb, ub = h.ub;
    // ---- mutation + clip (:144-155); Cr = 1 => trial = donor (:157) -------------------------------------------------
    if (tid < D) {
        const int d = tid;
        const double F = 0.5;
        auto x = [&](int row) { return (double)X[(size_t)row * D + d]; };
        const int r0 = mi[0], r1 = mi[1], r2 = mi[2], r3 = mi[3], r4 = mi[4];
        double v;
        if (a == 0) v = x(r0) + F * (x(r1) - x(r2));
        else if (a == 1) v = x(r0) + F * (x(r1) - x(r2) + x(r3) - x(r4));
        else if (a == 2) v = x(r0) + F * ((double)Xg[d] - x(r0) + x(r1) - x(r2) + x(r3) - x(r4));
        else v = x(p) + F * (x(r0) - x(p) + x(r1) - x(r2));
        v = fmin(fmax(v, lb), ub);
        const float t32 = (float)v;
        trial[d] = t32;
        if (dbg) dbg[(size_t)b * (D + 1) + d] = t32;
    }
    __syncthreads();
    PROF_MARK(1);
    // ---- median of the (pre-selection) costs (:172), then the fitness of the trial with the whole CTA (:159-163) ----------
    for (int i = tid; i < NP; i += MBX_NT) {
        const int rk = dd_rank(sc, i, NP);
        if (rk == (NP - 1) / 2) mf[1] = sc[i];
        if (rk == NP / 2) mf[2] = sc[i];
    }
    __syncthreads();
    PROF_MARK(2);
    {
        const double f = cta_eval_one<D, KIND>(h, cC, consts + gi->off, trial, scratch, ph, 0u, (uint32_t)fes0, lane,
                                              cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 : nullptr, 1);
        if (tid == 0) mf[0] = (float)f;
    }
    __syncthreads();
    PROF_MARK(3);
    const float tc = mf[0];
    const float cp = sc[p];
    const double dtc = (double)tc;
    // ---- reward, credit records (threads 1..4: one offspring metric each; their read-modify-writes of the history run in
    //      parallel with thread 0's window bookkeeping), window, stagnation, selection flags (:165-201) ---------------------
    const double med = ((double)mf[1] + (double)mf[2]) / 2.;              // np.median: mean of the middle pair
    const double om[4] = {(double)cp - dtc, (double)cpb - dtc, (double)cgb - dtc, med - dtc};
    if (tid >= 1 && tid <= 4) {
        const int m = tid - 1;
        const double omv = m == 0 ? om[0] : (m == 1 ? om[1] : (m == 2 ? om[2] : om[3]));
        if (omv > 0.) {
            const size_t k = (((size_t)b * 4 + a) * 4 + m) * DD_GEN_MAX;
            const int n = S.n_succ[k];
            S.om_sum[k] = n ? S.om_sum[k] + omv : omv;
            S.om_max[k] = n ? fmax(S.om_max[k], omv) : omv;
            S.n_succ[k] = n + 1;
        }
    }
    if (tid == 5) S.n_tot[((size_t)b * 4 + a) * DD_GEN_MAX] += 1;
    if (tid == 0) {
        if (dbg) dbg[(size_t)b * (D + 1) + D] = tc;
        reward[b] = (float)fmax((double)cp - dtc, 0.);
        // window OM_W (:176-186): when full drop the oldest entry of the same operator, else the entry with the largest
        // trial cost (first in list order on ties); the new entry goes to the end of the list
        int slot = wlen;
        if (wlen >= DD_W) {
            int k = -1; double ks = 1e300;
            for (int w = 0; w < wlen; ++w)
                if ((int)swin[w * DD_WS] == a && swin[w * DD_WS + 6] < ks) { ks = swin[w * DD_WS + 6]; k = w; }
            if (k < 0) {
                double bc = 0.; ks = 1e300;
                for (int w = 0; w < wlen; ++w) {
                    const double c = swin[w * DD_WS + 5], sq = swin[w * DD_WS + 6];
                    if (k < 0 || c > bc || (c == bc && sq < ks)) { bc = c; ks = sq; k = w; }
                }
            }
            slot = k;
        } else {
            wlen += 1;
        }
        double* e = swin + slot * DD_WS;
        e[0] = (double)a; e[1] = om[0]; e[2] = om[1]; e[3] = om[2]; e[4] = om[3]; e[5] = dtc; e[6] = (double)wseq; e[7] = 0.;
        double* ge = S.win + ((size_t)b * DD_W + slot) * DD_WS;
#pragma unroll
        for (int q = 0; q < DD_WS; ++q) ge[q] = e[q];
        wseq += 1;
        mi[12] = wlen;
        if (tc >= cgb) stag += 1;                                          // :188-189
        mi[13] = stag;
        const int replaced = (tc <= cp) ? 1 : 0;                           // :192
        const int newg = (replaced && tc <= cgb) ? 1 : 0;                  // :196
        mi[10] = replaced; mi[11] = newg;
    }
    __syncthreads();
    PROF_MARK(4);
    const int replaced = mi[10], newg = mi[11];
    wlen = mi[12];
    if (replaced) {
        if (tid < D) X[(size_t)p * D + tid] = trial[tid];                  // in place: a prebest view of row p follows (:194)
        if (tid == 0) { sc[p] = tc; S.cost[pb + p] = tc; }
        if (newg) {
            if (tid < D) Xg[tid] = trial[tid];
            cgb = tc; rebound = 1;
        }
    }
    if (tc > cgw) cgw = tc;                                                // :200-201
    const int p2 = (p + 1) % NP;
    const int fes = fes0 + 1;
    const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)cgb <= 1e-8) : (fes >= cfg.maxFEs);
    // next 5 indices (:85)
    if (tid < 5) {
        int rj;
        if (cfg.use_tape) rj = T.r5[(size_t)b * 5 + tid];
        else { uint4 w = ph((uint32_t)(tid >> 2), (uint32_t)fes0, lane, MBX_TAG_DDQN); rj = bounded(tid == 0 || tid == 4 ? w.x : (tid == 1 ? w.y : (tid == 2 ? w.z : w.w)), NP); }
        mi[5 + tid] = rj; S.r[(size_t)b * 5 + tid] = rj;
    }
    __syncthreads();
    stag = mi[13];
    if (tid == 0) {
        ints[DD_I_POINTER] = p2; ints[DD_I_GEN] = gen; ints[DD_I_STAG] = stag; ints[DD_I_PROW] = prow;
        ints[DD_I_REBOUND] = rebound; ints[DD_I_WLEN] = wlen; ints[DD_I_WSEQ] = wseq;
        S.c3[(size_t)b * 3] = cgb; S.c3[(size_t)b * 3 + 1] = cgw;
        int lg = S.log_index[b], clen = S.curve_len[b];
        float* curve = S.curve + (size_t)b * MBX_CURVE_LEN;
        if (fes >= lg * cfg.log_interval) {                                // :204-206
            lg += 1;
            if (clen < MBX_CURVE_LEN) curve[clen] = cgb;
            clen = min(clen + 1, MBX_CURVE_LEN);
        }
        if (done) {                                                        // :215-219
            if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = cgb;
            else { curve[clen] = cgb; clen += 1; }
        }
        S.log_index[b] = lg; S.curve_len[b] = clen; S.fes[b] = fes; S.done[b] = done ? 1 : 0;
    }
    PROF_MARK(5);
    dd_features<D>(cfg, S, b, sc, red, swin, p2, gen, stag, prow, wlen, fes, cgb, cgw, cpb, mi + 5, lb, ub,
                   feature + (size_t)b * DD_NFEAT);
    PROF_MARK(6);
}

// =====================================================================================================
template <int D, int KIND>
static int dd_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxDdqnState* s, const MbxDdqnTape* tape, float* feature, int B, cudaStream_t st) {
    auto kern = deddqn_init_kernel<D, KIND>;
    static bool attr_done = false;
    if (!attr_done) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PopSmem<D, KIND>::kBytes));
        attr_done = true;
    }
    MbxDdqnTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, PopSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, feature);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int dd_launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxDdqnState* s, const MbxDdqnTape* tape, const int32_t* action, float* feature,
                            float* reward, float* dbg, int B, cudaStream_t st) {
    auto kern = deddqn_step_kernel<D, KIND>;
    static bool attr_done = false;
    if (!attr_done) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DdqnSmem<D, KIND>::kBytes));
        attr_done = true;
    }
    MbxDdqnTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, DdqnSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, feature, reward, dbg);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    if (cfg->dim == 10 && kind == MBX_KIND_BBOB) return FN<10, MBX_KIND_BBOB>(__VA_ARGS__);           \
    if (cfg->dim == 30 && kind == MBX_KIND_BBOB) return FN<30, MBX_KIND_BBOB>(__VA_ARGS__);           \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_deddqn_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxDdqnState* s, const MbxDdqnTape* tape, float* feature, int B,
                           cudaStream_t st) {
    MBX_DISPATCH(dd_launch_init_t, inst, consts, lane_inst, cfg, s, tape, feature, B, st)
    mbx_set_error("mbx_deddqn_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_deddqn_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxDdqnState* s, const MbxDdqnTape* tape, const int32_t* action,
                           float* feature, float* reward, float* dbg, int B, cudaStream_t st) {
    MBX_DISPATCH(dd_launch_step_t, inst, consts, lane_inst, cfg, s, tape, action, feature, reward, dbg, B, st)
    mbx_set_error("mbx_deddqn_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
