// mbx_deddqn.cu -- fused DE-DDQN step (DE family, ONE individual per step): generation bookkeeping ->
// mutation by the chosen operator (rand/1, rand/2, rand-to-best/2, cur-to-rand/1; F = 0.5, Cr = 1) ->
// clip -> fitness of the single trial -> reward -> operator-credit history (N_tot / N_succ / OM per
// generation slot, window OM_W) -> greedy `<=` replacement, gbest / gworst / stagnation -> cost-curve
// log -> the 99 state features of the next individual.  One CTA per lane; the single trial is evaluated by the
// whole CTA (one dimension per thread, mbx_bbob_cta.cuh).
// Reference: optimizer/de_ddqn_optimizer.py:46-74 (init_population), :76-129 (__get_state), :131-220
// (update); operators/mutate.py:36-39,53-56,124-127,178-181; crossover.py:6-18; boundary_control.py:5-9.
// numpy twin: oracle/deddqn.py (documents the reference quirks that are reproduced here).
//
// This header holds the device code of the step: `dd_step_core` works on a lane context (`DdLane`: where the lane's state lives) and a
// register copy of its scalars (`DdRegs`), so that the SAME arithmetic serves
//   * the per-step kernel `deddqn_step_kernel` (mbx_deddqn.cu): state in HBM, one launch per function evaluation, action from outside;
//   * the persistent kernel `deddqn_run_kernel` (mbx_deddqn_run.cu): state resident in shared memory, the Q-network evaluated
//     in the kernel, n_steps function evaluations per launch.
#pragma once
#include "mbx_eval.cuh"
#include "mbx_launch.h"

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

// Sum of val(vt) over the 128 "virtual threads" vt of a lane, in ONE fixed order whatever the real group size: a xor-butterfly inside each
// virtual warp (all lanes of a butterfly end with the same bits: IEEE addition is commutative), then the four warp sums added in ascending
// order.  A group of MBX_NT < 128 threads walks 128 / MBX_NT virtual threads per thread -- same tree, same bits as the 128-thread kernel.
template <class F>
__device__ __forceinline__ double dd_block_sum(F val, double* red) {
    const int tid = MBX_TID, l = tid & 31;
    double part[128 / MBX_NT];
#pragma unroll
    for (int j = 0; j < 128 / MBX_NT; ++j) {
        double v = val(tid + j * MBX_NT);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        part[j] = v;
    }
    MBX_SYNC();
    if (l == 0) {
#pragma unroll
        for (int j = 0; j < 128 / MBX_NT; ++j) red[(tid >> 5) + j * (MBX_NT / 32)] = part[j];
    }
    MBX_SYNC();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) s += red[k];
    return s;
}
// sum of two per-thread counters over the lane group (warp reduction, then the warps through `scratch` [2 * MBX_NT / 32] when there are several)
__device__ __forceinline__ void dd_group_count2(int& a, int& b, int* scratch) {
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
#if MBX_NT > 32
    const int tid = MBX_TID;
    if ((tid & 31) == 0) { scratch[tid >> 5] = a; scratch[MBX_NT / 32 + (tid >> 5)] = b; }
    MBX_SYNC();
    a = 0; b = 0;
#pragma unroll
    for (int k = 0; k < MBX_NT / 32; ++k) { a += scratch[k]; b += scratch[MBX_NT / 32 + k]; }
#endif
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

// Where ONE lane's state lives.  Per-step kernel: X / history / X_gbest / X_prebest point into HBM, `sc` and `swin` are this launch's
// shared-memory copies of the costs and the window and every change is written through (`cost_wt`, `win_wt`, `r_wt`).  Persistent kernel:
// everything points into shared memory, the write-through pointers are null and the state goes back to HBM when the launch ends.
struct DdLane {
    float *X, *Xg, *Xp;                          // [NP,D], [D], [D]
    int32_t *n_tot, *n_succ;                     // [4,10], [4,4,10]
    double *om_sum, *om_max;                     // [4,4,10]
    float* curve;                                // [MBX_CURVE_LEN] (HBM in both kernels: touched at log points only)
    float* cost_wt; double* win_wt; int32_t* r_wt;      // write-through targets or nullptr
    float* sc; double* swin;                     // working copies in shared memory: costs [NP], window [W*WS]
    float* sorted;                               // persistent kernel: the costs in ascending order (NaN last), maintained incrementally; else null
    double *red, *scratch;                       // reduction scratch [16], evaluation scratch
    float *trial, *mf;                           // trial vector [D]; mf[0] trial cost, mf[1], mf[2] median pair
    int* mi;                                     // [0..4] r (old), [5..9] r (new), [10] replaced, [11] new gbest, [12] wlen, [13] stag
    float* feat;                                 // the 99 features of the next individual (HBM or shared memory)
};
// The lane's scalars, identical in all threads of the lane except `log_index`, `curve_len` (maintained by thread 0 only).
struct DdRegs {
    int p, gen, stag, prow, rebound, wlen, wseq, fes, log_index, curve_len;
    float cgb, cgw, cpb;
    bool done;
};

// ---- __get_state (de_ddqn_optimizer.py:76-129): all threads call; needs sc[] = current costs in smem ------------------
template <int D>
__device__ void dd_features(const MbxRunConfig& cfg, const DdLane& L, int p, int gen, int stag, int prow, int wlen, int fes, float cgb,
                            float cgw, float cpb, const int* r, double lb, double ub) {
    const int NP = cfg.NP, tid = MBX_TID;
    const float* sc = L.sc;
    double *red = L.red, *swin = L.swin;
    float* feat = L.feat;
    const double gb = (double)cgb, span = (double)cgw - (double)cgb;
    // mean / std of the costs (np.mean, np.std)
    const double mean = dd_block_sum([&](const int vt) { return vt < NP ? (double)sc[vt] : 0.; }, red) / (double)NP;
    const double var = dd_block_sum([&](const int vt) { const double dv = vt < NP ? ((double)sc[vt] - mean) : 0.; return dv * dv; }, red) / (double)NP;
    const float* X = L.X;
    const double cp = (double)sc[p];
    const double max_dist = sqrt((double)D * ((ub - lb) * (ub - lb)));
    // scalar features 0..5, 12..17: one lane each of warp 0, (numerator, denominator) picked by selects and ONE float64 division executed for
    // all of them at once (a division is a ~30-instruction sequence: per-feature branches would run it eleven times per warp)
    if (tid < 32) {
        const double sd = sqrt(var), fmx = (double)cfg.maxFEs;
        double num = 0., den = 1.;
        int idx = -1;
        if (tid == 0) { num = cp - gb; den = span; idx = 0; }
        if (tid == 1) { num = mean - gb; den = span; idx = 1; }
        if (tid == 2) { num = sd; den = span / 2.; idx = 2; }
        if (tid == 3) { num = (double)(cfg.maxFEs - fes); den = fmx; idx = 3; }
        if (tid == 4) { num = (double)stag; den = fmx; idx = 5; }
        if (tid == 5) { num = cp - (double)cpb; den = span; idx = 17; }
        if (tid >= 8 && tid < 13) { num = cp - (double)sc[r[tid - 8]]; den = span; idx = 12 + tid - 8; }
        const double q = num / den;
        if (idx >= 0) feat[idx] = (float)q;
        if (tid == 3) feat[4] = 1.f;
    }
    // 7 distances: X[p] to X[r_j] (6..10), X_prebest (11), X_gbest (18); warp w handles targets w, w + #warps, ...  The targets of a warp
    // are unrolled so that their row loads (HBM / L2 in the persistent kernel) are in flight together.
    {
        constexpr int NW = MBX_NT / 32, TPW = (7 + NW - 1) / NW;
        const int w = tid >> 5, l = tid & 31;
        double acc[TPW];
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
            const int t = w + j * NW;
            acc[j] = 0.;
            if (t < 7) {
                const float* Y = t < 5 ? X + (size_t)r[t] * D : (t == 5 ? (prow >= 0 ? X + (size_t)prow * D : L.Xp) : L.Xg);
                for (int d = l; d < D; d += 32) { const double df = (double)X[(size_t)p * D + d] - (double)Y[d]; acc[j] += df * df; }
            }
        }
        double sel = 0.;                                                   // lane j of the warp finishes its j-th target
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
            double a = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (l == j) sel = a;
        }
        const int t = w + l * NW;
        const double dist = sqrt(sel) / max_dist;                          // one sqrt + one division sequence for all targets of the warp
        if (l < TPW && t < 7) feat[t < 5 ? 6 + t : (t == 5 ? 11 : 18)] = (float)dist;
    }
    // operator-credit aggregates (features 19..98): item (op, m) = tid / 8, its 10 generation slots and 50 window entries spread over the 8
    // threads of the item (the float64 divisions k / n and the window scan were a serial chain of one thread per item), partial sums
    // combined by shuffles inside the aligned 8-lane group
#pragma unroll
    for (int slot = tid; slot < 128; slot += MBX_NT) {                     // 16 (operator, metric) items x 8 parts; one pass at 128 threads
        const int item = slot >> 3, part = slot & 7, op = item >> 2, m = item & 3;
        const int G = min(DD_GEN_MAX, gen);
        const int32_t* nt = L.n_tot + op * DD_GEN_MAX;
        const int32_t* ns = L.n_succ + (op * 4 + m) * DD_GEN_MAX;
        const double* os = L.om_sum + (op * 4 + m) * DD_GEN_MAX;
        const double* ox = L.om_max + (op * 4 + m) * DD_GEN_MAX;
        double f19 = 0., f35 = 0., f67 = 0., f83 = 0.;
        int tot = 0;
        for (int g = part; g < G; g += 8) {
            const int n = nt[g], k = ns[g];
            tot += n;
            if (n > 0) f19 += (double)k / (double)n;
            if (k > 0) { f35 += os[g]; f67 += ox[g]; }
        }
        // window sums (features 83..98): entries whose action == op (unordered slots: the order of the sum differs from the reference's
        // list order by rounding only)
        for (int w = part; w < wlen; w += 8)
            if ((int)swin[w * DD_WS] == op) f83 += swin[w * DD_WS + 1 + m];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            f19 += __shfl_xor_sync(0xffffffffu, f19, o); f35 += __shfl_xor_sync(0xffffffffu, f35, o);
            f67 += __shfl_xor_sync(0xffffffffu, f67, o); f83 += __shfl_xor_sync(0xffffffffu, f83, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if (part == 0) {
            if (tot > 0) f35 = f35 / (double)tot;
            double f51 = 0.;
            if (gen >= 2) {
                const int dn = nt[0] - nt[1];
                if (dn != 0 && ns[0] > 0 && ns[1] > 0) f51 = (ox[0] - ox[1]) / (ox[1] * fabs((double)dn));
            }
            feat[19 + item] = (float)f19; feat[35 + item] = (float)f35; feat[51 + item] = (float)f51;
            feat[67 + item] = (float)f67; feat[83 + item] = (float)f83;
        }
    }
}

// total order on costs with NaN last (the order dd_rank sorts by)
__device__ __forceinline__ bool dd_lt(const float a, const float b) { return a < b || (a == a && b != b); }

// ---- one environment step of one lane (de_ddqn_optimizer.py:131-220 + __get_state :76-129).  ALL threads of the lane call.  `a` in 0..3;
//      `h` = the lane's instance header, `cC` = its constants (shared memory for protein, global / L1 for BBOB), `gC` = the constants in
//      global memory.  Returns the reward max(parent - trial, 0) (valid in every thread). -------------------------------------------
template <int D, int KIND>
__device__ __forceinline__ float dd_step_core(const MbxRunConfig& cfg, const MbxInstance& h, const double* cC, const double* gC,
                                              const DdLane& L, DdRegs& R, const int a, const Philox& ph, const uint32_t lane,
                                              const double* noise_tape, const int32_t* r5_tape, float* dbg) {
    PROF_INIT
    const int NP = cfg.NP, tid = MBX_TID;
    float* X = L.X;
    float *Xg = L.Xg, *Xp = L.Xp;
    float* sc = L.sc;
    double* swin = L.swin;
    float* trial = L.trial;
    float* mf = L.mf;
    int* mi = L.mi;
    const int p = R.p;
    int gen = R.gen, stag = R.stag, prow = R.prow, rebound = R.rebound, wlen = R.wlen, wseq = R.wseq;
    const int fes0 = R.fes;
    float cgb = R.cgb, cgw = R.cgw;
    const float cpb = R.cpb;                                               // never refreshed (:135)
    // ---- generation boundary (:132-142) ------------------------------------------------------------------------
    if (p == 0) {
        if (rebound) {                                                     // X_prebest = X_gbest (a trial array by now)
            prow = -1;
            for (int d = tid; d < D; d += MBX_NT) Xp[d] = Xg[d];
        }
        gen += 1;
        // appendleft(0) on the maxlen-10 deques: slot g <- slot g-1
        int32_t* nt = L.n_tot;
        int32_t* ns = L.n_succ;
        double* os = L.om_sum;
        double* ox = L.om_max;
        constexpr int Q = (16 * DD_GEN_MAX + MBX_NT - 1) / MBX_NT;          // 2 values per thread at 128 threads
        int vt[Q], vs[Q]; double vo[Q], vx[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = tid + q * MBX_NT;
            vt[q] = 0; vs[q] = 0; vo[q] = 0.; vx[q] = 0.;
            if (k < 4 * DD_GEN_MAX && (k % DD_GEN_MAX)) vt[q] = nt[k - 1];
            if (k < 16 * DD_GEN_MAX && (k % DD_GEN_MAX)) { vs[q] = ns[k - 1]; vo[q] = os[k - 1]; vx[q] = ox[k - 1]; }
        }
        MBX_SYNC();
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = tid + q * MBX_NT;
            if (k < 4 * DD_GEN_MAX) nt[k] = vt[q];
            if (k < 16 * DD_GEN_MAX) { ns[k] = vs[q]; os[k] = vo[q]; ox[k] = vx[q]; }
        }
    }
    MBX_SYNC();
    PROF_MARK(0);
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
        v = v < lb ? lb : v; v = v > ub ? ub : v;               // np.clip (keeps a NaN, unlike fmin / fmax)
        const float t32 = (float)v;
        trial[d] = t32;
        if (dbg) dbg[d] = t32;
    }
    MBX_SYNC();
    PROF_MARK(1);
    // ---- median of the (pre-selection) costs (:172), then the fitness of the trial with the whole lane group (:159-163) ----------
    // The per-step kernel ranks all costs (NP^2 comparisons); the persistent kernel keeps an ascending copy of the costs that changes by
    // one remove + one insert per step (below), so the middle pair is read directly.  Same two values either way.
    if (L.sorted) {
        if (tid == 0) { mf[1] = L.sorted[(NP - 1) / 2]; mf[2] = L.sorted[NP / 2]; }
    } else {
        for (int i = tid; i < NP; i += MBX_NT) {
            const int rk = dd_rank(sc, i, NP);
            if (rk == (NP - 1) / 2) mf[1] = sc[i];
            if (rk == NP / 2) mf[2] = sc[i];
        }
        MBX_SYNC();
    }
    PROF_MARK(2);
    {
        const double f = cta_eval_one<D, KIND>(h, cC, gC, trial, L.scratch, ph, 0u, (uint32_t)fes0, lane, noise_tape, 1);
        if (tid == 0) mf[0] = (float)f;
    }
    MBX_SYNC();
    PROF_MARK(3);
    const float tc = mf[0];
    const float cp = sc[p];
    const double dtc = (double)tc;
    // ---- reward, credit records (threads 1..4: one offspring metric each; their read-modify-writes of the history run in
    //      parallel with thread 0's window bookkeeping), window, stagnation, selection flags (:165-201) ---------------------
    const double med = ((double)mf[1] + (double)mf[2]) / 2.;              // np.median: mean of the middle pair
    const double om[4] = {(double)cp - dtc, (double)cpb - dtc, (double)cgb - dtc, med - dtc};
    const float rew = (float)fmax((double)cp - dtc, 0.);
    if (tid >= 1 && tid <= 4) {
        const int m = tid - 1;
        const double omv = m == 0 ? om[0] : (m == 1 ? om[1] : (m == 2 ? om[2] : om[3]));
        if (omv > 0.) {
            const int k = (a * 4 + m) * DD_GEN_MAX;
            const int n = L.n_succ[k];
            L.om_sum[k] = n ? L.om_sum[k] + omv : omv;
            L.om_max[k] = n ? fmax(L.om_max[k], omv) : omv;
            L.n_succ[k] = n + 1;
        }
    }
    if (tid == 5) L.n_tot[a * DD_GEN_MAX] += 1;
    if (tid < 32) {
        // window OM_W (:176-186): when full drop the oldest entry of the same operator, else the entry with the largest trial cost
        // (np.argmax: the first in list order on ties, a NaN counts as the maximum); the new entry goes to the end of the list.
        // Warp 0 scans the 50 slots two per lane and reduces by shuffles.
        int slot = wlen;
        if (wlen >= DD_W) {
            double ks = 1e300; int k = -1;
            for (int w = tid; w < wlen; w += 32)
                if ((int)swin[w * DD_WS] == a && swin[w * DD_WS + 6] < ks) { ks = swin[w * DD_WS + 6]; k = w; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oks = __shfl_xor_sync(0xffffffffu, ks, o);
                const int ok = __shfl_xor_sync(0xffffffffu, k, o);
                if (oks < ks) { ks = oks; k = ok; }                        // sequence numbers are distinct
            }
            if (k < 0) {
                double bc = 0.; ks = 1e300;
                for (int w = tid; w < wlen; w += 32) {
                    const double c = swin[w * DD_WS + 5], sq = swin[w * DD_WS + 6];
                    const bool gt = k < 0 || (c != c ? (bc == bc || sq < ks) : (bc == bc && (c > bc || (c == bc && sq < ks))));
                    if (gt) { bc = c; ks = sq; k = w; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double obc = __shfl_xor_sync(0xffffffffu, bc, o), oks = __shfl_xor_sync(0xffffffffu, ks, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, k, o);
                    const bool gt = ok >= 0 && (k < 0 || (obc != obc ? (bc == bc || oks < ks) : (bc == bc && (obc > bc || (obc == bc && oks < ks)))));
                    if (gt) { bc = obc; ks = oks; k = ok; }
                }
            }
            slot = k;
        } else {
            wlen += 1;
        }
        if (tid < DD_WS) {                                                 // lane q writes field q of the new entry
            const double ev = tid == 0 ? (double)a : (tid <= 4 ? (tid == 1 ? om[0] : (tid == 2 ? om[1] : (tid == 3 ? om[2] : om[3])))
                                                                : (tid == 5 ? dtc : (tid == 6 ? (double)wseq : 0.)));
            swin[slot * DD_WS + tid] = ev;
            if (L.win_wt) L.win_wt[(size_t)slot * DD_WS + tid] = ev;
        }
        if (tid == 0) {
            if (dbg) dbg[D] = tc;
            mi[12] = wlen;
            if (tc >= cgb) stag += 1;                                      // :188-189
            mi[13] = stag;
            const int replaced = (tc <= cp) ? 1 : 0;                       // :192
            const int newg = (replaced && tc <= cgb) ? 1 : 0;              // :196
            mi[10] = replaced; mi[11] = newg;
        }
    }
    wseq += 1;
    MBX_SYNC();
    PROF_MARK(4);
    const int replaced = mi[10], newg = mi[11];
    wlen = mi[12];
    if (replaced) {
        if (tid < D) X[(size_t)p * D + tid] = trial[tid];                  // in place: a prebest view of row p follows (:194)
        if (tid == 0) { sc[p] = tc; if (L.cost_wt) L.cost_wt[p] = tc; }
        if (newg) {
            if (tid < D) Xg[tid] = trial[tid];
            cgb = tc; rebound = 1;
        }
    }
    if (L.sorted && replaced) {
        // the ascending copy: remove the first occurrence of the parent's cost cp, insert tc behind its equals.  r = #elements < cp is
        // that occurrence's position; c = #elements <= tc; every thread reads the neighbourhoods of its elements before the barrier and
        // writes after it.
        float* ss = L.sorted;
        constexpr int E = MBX_MAX_NP / MBX_NT;                             // elements per thread
        float vm[E], v0[E], vp[E];
        int cr = 0, cq = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int k = tid + j * MBX_NT;
            vm[j] = (k > 0 && k < NP) ? ss[k - 1] : 0.f; v0[j] = k < NP ? ss[k] : 0.f; vp[j] = k + 1 < NP ? ss[k + 1] : 0.f;
            cr += (k < NP && dd_lt(v0[j], cp)) ? 1 : 0;
            cq += (k < NP && !dd_lt(tc, v0[j])) ? 1 : 0;
        }
        dd_group_count2(cr, cq, mi + 16);
        MBX_SYNC();
        const int r = cr, q = cq - (dd_lt(tc, cp) ? 0 : 1);                // the removed element is not counted
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int k = tid + j * MBX_NT;
            if (k < NP) {
                float val = tc;
                if (k != q) {
                    const int jj = k < q ? k : k - 1;                      // index in the array without the removed element
                    val = jj < r ? (jj == k ? v0[j] : vm[j]) : (jj + 1 == k ? v0[j] : vp[j]);
                }
                ss[k] = val;
            }
        }
    }
    if (tc > cgw) cgw = tc;                                                // :200-201
    const int p2 = (p + 1) % NP;
    const int fes = fes0 + 1;
    const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)cgb <= 1e-8) : (fes >= cfg.maxFEs);
    // next 5 indices (:85)
    if (tid < 5) {
        int rj;
        if (r5_tape) rj = r5_tape[tid];
        else { uint4 w = ph((uint32_t)(tid >> 2), (uint32_t)fes0, lane, MBX_TAG_DDQN); rj = bounded(tid == 0 || tid == 4 ? w.x : (tid == 1 ? w.y : (tid == 2 ? w.z : w.w)), NP); }
        mi[5 + tid] = rj;
        if (L.r_wt) L.r_wt[tid] = rj;
    }
    MBX_SYNC();
    stag = mi[13];
    if (tid == 0) {
        int lg = R.log_index, clen = R.curve_len;
        float* curve = L.curve;
        if (fes >= lg * cfg.log_interval) {                                // :204-206
            lg += 1;
            if (clen < MBX_CURVE_LEN) curve[clen] = cgb;
            clen = min(clen + 1, MBX_CURVE_LEN);
        }
        if (done) {                                                        // :215-219
            if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = cgb;
            else { curve[clen] = cgb; clen += 1; }
        }
        R.log_index = lg; R.curve_len = clen;
    }
    R.p = p2; R.gen = gen; R.stag = stag; R.prow = prow; R.rebound = rebound; R.wlen = wlen; R.wseq = wseq; R.fes = fes;
    R.cgb = cgb; R.cgw = cgw; R.done = done;
    PROF_MARK(5);
    dd_features<D>(cfg, L, p2, gen, stag, prow, wlen, fes, cgb, cgw, cpb, mi + 5, lb, ub);
    PROF_MARK(6);
    return rew;
}
