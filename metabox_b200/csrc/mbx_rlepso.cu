// mbx_rlepso.cu -- fused RLEPSO generation (PSO family): velocity components (CLPSO tournament, FDR
// argmin, gbest, pbest) -> clip -> position -> fitness -> pbest/gbest -> patience -> stochastic
// re-initialisation -> cost-curve log, one CTA per lane, state float32 in HBM / float64 in registers.
// Reference: optimizer/rlepso_optimizer.py:39-65 (init_population), :76-132 (velocity parts,
// coefficients), :134-168 (re-init), :173-263 (update).  numpy twin: oracle/rlepso.py.
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#define FDR_IB 2      // ranks per block in the sorted FDR scan

// Shared-memory plan, sized by the run-time population size (NP = 100 -> 62 KB at D = 30: 3 CTAs per SM).
struct RlLayout {
    int NP, NPp;
    size_t oConst, oA, oPer, oGco, oGx, oXs, oF, oI, oTgt, oRed, bytes;
    __host__ __device__ RlLayout(int np, int D, int kConst, bool protein) {
        NP = np; NPp = (np + 3) & ~3;
        const size_t scratch = protein ? (size_t)(MBX_NT / 32) * MBX_PROT_WARP_SCRATCH : (size_t)D * NPp;
        const size_t regionA = scratch > (size_t)NPp * D ? scratch : (size_t)NPp * D;
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oA = oConst + sizeof(double) * kConst;               // sorted pbest_x as f64 [NP*D]  /  eval scratch
        oPer = oA + sizeof(double) * regionA;                // 5 per-particle double arrays
        oGco = oPer + sizeof(double) * 5 * NPp;              // 5 groups x (w, c1..c4)
        oGx = oGco + sizeof(double) * 32;                    // gbest_x as double [D]
        oXs = oGx + sizeof(double) * MBX_MAX_DIM;            // new positions f32 [NP*D]
        oF = oXs + sizeof(float) * NPp * D;                  // 4 per-particle float arrays
        oI = oF + sizeof(float) * 4 * NPp;                   // 3 per-particle int arrays
        oTgt = oI + sizeof(int) * 3 * NPp;                   // FDR targets u8 [NP*D] + 2 masks u8 [NP]
        oRed = (oTgt + (size_t)NPp * D + 2 * NPp + 15) & ~(size_t)15;
        bytes = oRed + 64;
    }
};

template <int D, int KIND>
struct RlepsoViews {
    MbxInstance* h;
    double *sC, *A, *spf, *cm, *r1, *r2, *ur, *gco, *gx;
    float *xs, *pf32, *nc, *cc, *tmpf;
    int *pni, *rank, *order;
    unsigned char *tgt, *mask, *better;
    float* redv;
    int* redi;
    __device__ __forceinline__ RlepsoViews(unsigned char* s, const RlLayout& L) {
        const int N = L.NPp;
        h = reinterpret_cast<MbxInstance*>(s);
        sC = reinterpret_cast<double*>(s + L.oConst);
        A = reinterpret_cast<double*>(s + L.oA);
        double* per = reinterpret_cast<double*>(s + L.oPer);
        spf = per; cm = per + N; r1 = per + 2 * N; r2 = per + 3 * N; ur = per + 4 * N;
        gco = reinterpret_cast<double*>(s + L.oGco);
        gx = reinterpret_cast<double*>(s + L.oGx);
        xs = reinterpret_cast<float*>(s + L.oXs);
        float* f = reinterpret_cast<float*>(s + L.oF);
        pf32 = f; nc = f + N; cc = f + 2 * N; tmpf = f + 3 * N;
        int* ii = reinterpret_cast<int*>(s + L.oI);
        pni = ii; rank = ii + N; order = ii + 2 * N;
        tgt = s + L.oTgt; mask = tgt + N * D; better = mask + N;
        redv = reinterpret_cast<float*>(s + L.oRed);
        redi = reinterpret_cast<int*>(s + L.oRed + 32);
    }
};

// cost-curve protocol shared by all optimizers (rlepso_optimizer.py:241-261)
__device__ __forceinline__ bool log_and_done(const MbxRunConfig& cfg, bool has_opt, int fes, float gbest, int& log_index,
                                             int& clen, float* curve) {
    if (fes >= log_index * cfg.log_interval) {
        log_index += 1;
        if (clen < MBX_CURVE_LEN) curve[clen] = gbest;
        clen = min(clen + 1, MBX_CURVE_LEN);
    }
    bool done = has_opt ? (fes >= cfg.maxFEs || (double)gbest <= 1e-8) : (fes >= cfg.maxFEs);
    if (done) {
        if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = gbest;
        else { curve[clen] = gbest; clen += 1; }
    }
    return done;
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
rlepso_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   MbxRunConfig cfg, MbxRlepsoState S, MbxRlepsoTape T, float* __restrict__ feature) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, NP = cfg.NP, tid = threadIdx.x;
    const RlLayout L(NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN);
    RlepsoViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);
    __syncthreads();
    const MbxInstance& h = *V.h;
    Philox ph(cfg.seed);
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb);
    const size_t base = (size_t)b * NP * D;
    for (int e = tid; e < NP * D; e += MBX_NT) {
        double up, uv;
        if (cfg.use_tape) { up = T.init_pos_u[base + e]; uv = T.init_vel_u[base + e]; }
        else { uint4 w = ph((uint32_t)e, MBX_GEN_INIT, (uint32_t)(cfg.lane_base + b), MBX_TAG_INIT); up = u01(w.x); uv = u01(w.y); }
        float p = (float)(lb + (ub - lb) * up);
        float v = (float)(-vmax + (vmax - (-vmax)) * uv);
        V.xs[e] = p;
        S.x[base + e] = p; S.pbest_x[base + e] = p; S.v[base + e] = v;
    }
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, V.nc, true, ph, MBX_GEN_INIT, (uint32_t)(cfg.lane_base + b),
                                         MBX_TAG_NOISE, cfg.use_tape && T.noise ? T.noise + (size_t)b * 6 * NP : nullptr, NP, nullptr, L.NPp);
    __syncthreads();
    float v = INFINITY; int ix = tid;
    if (tid < NP) {
        v = V.nc[tid];
        S.c_cost[(size_t)b * NP + tid] = v; S.pbest[(size_t)b * NP + tid] = v; S.pni[(size_t)b * NP + tid] = 0;
    }
    block_argmin<MBX_NT>(v, ix, V.redv, V.redi);
    for (int d = tid; d < D; d += MBX_NT) S.gbest_x[(size_t)b * D + d] = V.xs[ix * D + d];
    if (tid == 0) {
        S.gbest_val[b] = v; S.gbest_idx[b] = ix; S.no_improve[b] = 0; S.fes[b] = NP; S.log_index[b] = 1;
        S.curve[(size_t)b * MBX_CURVE_LEN] = v; S.curve_len[b] = 1; S.gen[b] = 0; S.done[b] = 0;
        feature[b] = (float)((double)NP / (double)cfg.maxFEs);
    }
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT, 3)
rlepso_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   MbxRunConfig cfg, MbxRlepsoState S, MbxRlepsoTape T, const double* __restrict__ pci,
                   const float* __restrict__ action, float* __restrict__ feature, float* __restrict__ reward,
                   int32_t* __restrict__ dbg) {
    const int b = blockIdx.x;
    if (S.done[b]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NP = cfg.NP, tid = threadIdx.x;
    const RlLayout L(NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN);
    RlepsoViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    const uint32_t gen = (uint32_t)S.gen[b];
    Philox ph(cfg.seed);
    double* spx = V.A;                                   // pbest_x as f64, rows in ascending-pbest order
    // ---- P0: per-particle scalars, draws, coefficients (:112-132) -------------------------------------------
    for (int d = tid; d < D; d += MBX_NT) V.gx[d] = (double)S.gbest_x[(size_t)b * D + d];
    if (tid < NP) {
        const int i = tid;
        V.pf32[i] = S.pbest[pb + i]; V.cc[i] = S.c_cost[pb + i];
        const int pni = S.pni[pb + i];
        V.pni[i] = pni;
        const int per = NP / 5;
        int g = i / per; if (g > 4) g = 4;
        const float a0 = action[(size_t)b * 35 + g * 5];             // stride 5, width 7 (overlapping windows)
        V.cm[i] = (double)(a0 * 0.01f) * (double)pni;
        if (cfg.use_tape) { V.r1[i] = T.rand1[pb + i]; V.r2[i] = T.rand2[pb + i]; V.ur[i] = T.reinit_u[pb + i]; }
        else { uint4 w4 = ph((uint32_t)i, gen, (uint32_t)(cfg.lane_base + b), MBX_TAG_P); V.r1[i] = u01(w4.x); V.r2[i] = u01(w4.y); V.ur[i] = u01(w4.z); }
    }
    if (tid < 5) {                                                   // float32 coefficient math (numpy-2 promotion)
        const float* a = action + (size_t)b * 35 + tid * 5;
        const float a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6];
        const float w = a1 * 0.8f + 0.1f;
        float ssum = a3 + a4; ssum = ssum + a5; ssum = ssum + a6; ssum = ssum + 1e-5f;
        float scale = 1.f / ssum; scale = scale * a2; scale = scale * 8.f;
        double* gc = V.gco + tid * 5;
        gc[0] = (double)w; gc[1] = (double)(scale * a3); gc[2] = (double)(scale * a4);
        gc[3] = (double)(scale * a5); gc[4] = (double)(scale * a6);
    }
    __syncthreads();
    // ---- P0b: stable rank of the pbest costs; sorted f64 copies of pbest / pbest_x ----------------------------------
    if (tid < NP) {
        const float fi = V.pf32[tid];
        int r = 0;
        for (int j = 0; j < NP; ++j) { const float fj = V.pf32[j]; r += (fj < fi || (fj == fi && j < tid)) ? 1 : 0; }
        V.rank[tid] = r; V.order[r] = tid; V.spf[r] = (double)fi;
    }
    __syncthreads();
    for (int e = tid; e < NP * D; e += MBX_NT) { const int i = e / D; spx[V.rank[i] * D + (e - i * D)] = (double)S.pbest_x[base + e]; }
    __syncthreads();
    const MbxInstance& h = *V.h;
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb);
    // ---- P1: FDR targets  argmin_j (pf[j]-pf[i]) / (|px[j,d]-px[i,d]|+1e-5)   (:97-103) ------------------
    // The minimum is <= 0 (j = i gives 0), so only particles with a strictly better pbest -- the ones BEFORE i in the
    // sorted order -- and exact ties can win: particle of rank r scans s < r only (half the pairs on average).
    // a/b < a'/b'  <=>  a*b' < a'*b  (b,b' > 0): cross-multiplication in f64; exact ties -> lowest original index (np.argmin).
    // Work item = (pair of rank blocks {m, nb-1-m}, d): every item scans ~NP candidates.
    {
        const int nb = (NP + FDR_IB - 1) / FDR_IB, npair = (nb + 1) / 2;
        for (int w = tid; w < npair * D; w += MBX_NT) {
            const int pq = w / D, d = w - pq * D;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int m = half ? nb - 1 - pq : pq;
                if (half && m == pq) break;
                const int ra = m * FDR_IB, rb = ra + 1;
                const bool hasb = rb < NP;
                const double pxa = spx[ra * D + d], pfa = V.spf[ra];
                const double pxb = hasb ? spx[rb * D + d] : pxa, pfb = hasb ? V.spf[rb] : pfa;
                double abA = 0., bbA = 1e-5, abB = 0., bbB = 1e-5;       // start from j = i (ratio 0)
                int jsA = ra, jsB = rb;
#pragma unroll 2
                for (int s = 0; s < ra; ++s) {
                    const double pxs = spx[s * D + d], pfs = V.spf[s];
                    const double a1 = pfs - pfa, q1 = fabs(pxs - pxa) + 1e-5;
                    const double a2 = pfs - pfb, q2 = fabs(pxs - pxb) + 1e-5;
                    const double l1 = a1 * bbA, r1 = abA * q1, l2 = a2 * bbB, r2 = abB * q2;
                    if (l1 < r1 || (l1 == r1 && V.order[s] < V.order[jsA])) { abA = a1; bbA = q1; jsA = s; }
                    if (l2 < r2 || (l2 == r2 && V.order[s] < V.order[jsB])) { abB = a2; bbB = q2; jsB = s; }
                }
                if (hasb) {                                              // candidate s = ra for the second rank of the block
                    const double a2 = pfa - pfb, q2 = fabs(pxa - pxb) + 1e-5;
                    const double l2 = a2 * bbB, r2 = abB * q2;
                    if (l2 < r2 || (l2 == r2 && V.order[ra] < V.order[jsB])) { abB = a2; bbB = q2; jsB = ra; }
                    V.tgt[V.order[rb] * D + d] = (unsigned char)jsB;
                }
                V.tgt[V.order[ra] * D + d] = (unsigned char)jsA;
            }
        }
    }
    __syncthreads();
    // ---- P2: velocity + position, element-wise over (i,d)   (:76-95, :105-109, :179-195) ---------------
    for (int e = tid; e < NP * D; e += MBX_NT) {
        const int i = e / D, d = e - i * D;
        double ru, fu; int t0, t1;
        if (cfg.use_tape) {
            ru = T.clpso_u[base + e]; fu = T.fdr_u[base + e];
            t0 = T.tour[(base + e) * 2]; t1 = T.tour[(base + e) * 2 + 1];
        } else {
            uint4 w = ph((uint32_t)e, gen, (uint32_t)(cfg.lane_base + b), MBX_TAG_PD);
            ru = u01(w.x); t0 = bounded(w.y, NP); t1 = bounded(w.z, NP); fu = u01(w.w);
        }
        const double x = (double)S.x[base + e], v = (double)S.v[base + e], pxi = spx[V.rank[i] * D + d];
        const int tsel = (V.pf32[t1] < V.pf32[t0]) ? t1 : t0;          // argmin of the 2 picks, tie -> first
        const double own = (ru > pci[i]) ? pxi : spx[V.rank[tsel] * D + d];
        const double v_clpso = ru * (own - x);
        const int tg = V.tgt[e];                                       // sorted position of the FDR target
        if (dbg) dbg[base + e] = V.order[tg];
        const double v_fdr = fu * (spx[tg * D + d] - pxi);
        const double v_pbest = V.r1[i] * (pxi - x);
        const double v_gbest = V.r2[i] * (V.gx[d] - x);
        const int per = NP / 5;
        int g = i / per; if (g > 4) g = 4;
        const double* gc = V.gco + g * 5;
        double nv = gc[0] * v + gc[1] * v_clpso + gc[2] * v_fdr + gc[3] * v_gbest + gc[4] * v_pbest;
        nv = fmin(fmax(nv, -vmax), vmax);
        double nx = fmin(fmax(x + nv, lb), ub);
        const float nx32 = (float)nx;
        V.xs[e] = nx32;
        S.x[base + e] = nx32; S.v[base + e] = (float)nv;
    }
    __syncthreads();
    // ---- P3: fitness of the NP new positions (:68-74, :198) ---------------------------------------------
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, V.nc, true, ph, gen, (uint32_t)(cfg.lane_base + b),
                                         MBX_TAG_NOISE, cfg.use_tape && T.noise ? T.noise + (size_t)b * 6 * NP : nullptr, NP, nullptr, L.NPp);
    __syncthreads();
    // ---- P4: pbest / gbest / patience / re-init mask (:200-238) ------------------------------------------
    int fes = S.fes[b] + NP;
    float gbest = S.gbest_val[b];
    const float pre_gbest = gbest;
    float v = INFINITY; int ix = tid;
    bool mymask = false;
    if (tid < NP) {
        const int i = tid;
        const float ncv = V.nc[i];
        v = ncv;
        const bool bt = ncv < V.pf32[i];
        V.better[i] = bt;
        if (bt) { V.pf32[i] = ncv; S.pbest[pb + i] = ncv; }
        S.c_cost[pb + i] = ncv;
        int pni = (ncv < V.cc[i]) ? 0 : V.pni[i] + 1;
        V.pni[i] = pni;
        V.cc[i] = ncv;
        mymask = V.ur[i] < V.cm[i] * 0.01 * (double)pni;
        V.mask[i] = mymask;
    }
    const int any_mask = __syncthreads_or(mymask ? 1 : 0);
    for (int e = tid; e < NP * D; e += MBX_NT)
        if (V.better[e / D]) S.pbest_x[base + e] = V.xs[e];
    block_argmin<MBX_NT>(v, ix, V.redv, V.redi);
    int gidx = -1;
    if (v < gbest) { gbest = v; gidx = ix; }
    if (gidx >= 0 && !any_mask)
        for (int d = tid; d < D; d += MBX_NT) S.gbest_x[(size_t)b * D + d] = V.xs[gidx * D + d];
    // ---- P5: stochastic re-initialisation (:134-168) -- rare ------------------------------------------------
    if (any_mask) {
        // gbest_x of the main pass must be captured before xs rows are overwritten
        if (gidx >= 0 && tid < D) V.gx[tid] = (double)V.xs[gidx * D + tid];
        __syncthreads();
        for (int e = tid; e < NP * D; e += MBX_NT) {
            const int i = e / D;
            if (!V.mask[i]) continue;
            double up, uv;
            if (cfg.use_tape) { up = T.reinit_pos_u[base + e]; uv = T.reinit_vel_u[base + e]; }
            else { uint4 w = ph((uint32_t)e, gen, (uint32_t)(cfg.lane_base + b), MBX_TAG_REINIT); up = u01(w.x); uv = u01(w.y); }
            const float p = (float)(lb + (ub - lb) * up);
            const float vv = (float)(-vmax + (vmax - (-vmax)) * uv);
            V.xs[e] = p; S.x[base + e] = p; S.v[base + e] = vv;
        }
        __syncthreads();
        // all NP are re-evaluated by the reference; for a noise-free instance the unmasked costs are
        // unchanged by construction, so only masked rows are recomputed there.
        const bool noisy = (h.noise != MBX_NOISE_NONE);
        float* c2 = V.tmpf;
        if (tid < NP) c2[tid] = V.nc[tid];
        __syncthreads();
        cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, c2, true, ph, gen, (uint32_t)(cfg.lane_base + b),
                                             MBX_TAG_NOISE2,
                                             cfg.use_tape && T.noise ? T.noise + (size_t)b * 6 * NP + 3 * NP : nullptr, NP,
                                             noisy ? nullptr : V.mask, L.NPp);
        __syncthreads();
        int nmask = 0;
        float v2 = INFINITY; int ix2 = tid;
        if (tid < NP) {
            const int i = tid;
            const float cv = c2[i];
            v2 = cv;
            const bool bt = cv < V.pf32[i];
            V.better[i] = bt;
            if (bt) { V.pf32[i] = cv; S.pbest[pb + i] = cv; }
            S.c_cost[pb + i] = cv;
        }
        nmask = __syncthreads_count(tid < NP && V.mask[tid]);
        fes += nmask;                                                   // only sum(mask) is charged (:141-143)
        for (int e = tid; e < NP * D; e += MBX_NT)
            if (V.better[e / D]) S.pbest_x[base + e] = V.xs[e];
        block_argmin<MBX_NT>(v2, ix2, V.redv, V.redi);
        if (v2 < gbest) {
            gbest = v2; gidx = ix2;
            for (int d = tid; d < D; d += MBX_NT) S.gbest_x[(size_t)b * D + d] = V.xs[ix2 * D + d];
        } else if (gidx >= 0) {
            for (int d = tid; d < D; d += MBX_NT) S.gbest_x[(size_t)b * D + d] = (float)V.gx[d];
        }
    }
    if (tid < NP) S.pni[pb + tid] = V.pni[tid];
    // ---- P6: bookkeeping (:225-228, :241-261) ----------------------------------------------------------------
    if (tid == 0) {
        if (gidx >= 0) { S.gbest_val[b] = gbest; S.gbest_idx[b] = gidx; }
        int lg = S.log_index[b], clen = S.curve_len[b];
        bool done = log_and_done(cfg, h.has_optimum != 0, fes, gbest, lg, clen, S.curve + (size_t)b * MBX_CURVE_LEN);
        S.log_index[b] = lg; S.curve_len[b] = clen; S.fes[b] = fes; S.gen[b] = (int)gen + 1;
        S.no_improve[b] = (gbest < pre_gbest) ? 0 : S.no_improve[b] + 1;
        S.done[b] = done ? 1 : 0;
        feature[b] = (float)((double)fes / (double)cfg.maxFEs);
        reward[b] = (gbest < pre_gbest) ? 1.f : -1.f;
    }
}

// =====================================================================================================
template <int D, int KIND>
static int launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                         const MbxRlepsoState* s, const MbxRlepsoTape* tape, float* feature, int B, cudaStream_t st) {
    auto kern = rlepso_init_kernel<D, KIND>;
    static bool attr_done = false;
    if (!attr_done) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    MbxRlepsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, RlLayout(cfg->NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, feature);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                         const MbxRlepsoState* s, const MbxRlepsoTape* tape, const double* pci, const float* action,
                         float* feature, float* reward, int32_t* dbg, int B, cudaStream_t st) {
    auto kern = rlepso_step_kernel<D, KIND>;
    static bool attr_done = false;
    if (!attr_done) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    MbxRlepsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, RlLayout(cfg->NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, pci, action, feature, reward, dbg);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    if (cfg->dim == 10 && kind == MBX_KIND_BBOB) return FN<10, MBX_KIND_BBOB>(__VA_ARGS__);           \
    if (cfg->dim == 30 && kind == MBX_KIND_BBOB) return FN<30, MBX_KIND_BBOB>(__VA_ARGS__);           \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_rlepso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlepsoState* s, const MbxRlepsoTape* tape, float* feature, int B,
                           cudaStream_t st) {
    MBX_DISPATCH(launch_init_t, inst, consts, lane_inst, cfg, s, tape, feature, B, st)
    mbx_set_error("mbx_rlepso_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_rlepso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlepsoState* s, const MbxRlepsoTape* tape, const double* pci,
                           const float* action, float* feature, float* reward, int32_t* dbg, int B, cudaStream_t st) {
    MBX_DISPATCH(launch_step_t, inst, consts, lane_inst, cfg, s, tape, pci, action, feature, reward, dbg, B, st)
    mbx_set_error("mbx_rlepso_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
