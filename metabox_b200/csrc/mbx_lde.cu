// mbx_lde.cu -- fused LDE generation (DE family, whole population per step): fitness sort ->
// min-max normalisation + numpy-exact 5-bin histogram -> current-to-pbest/1 mutation (+ donor
// difference) -> binomial crossover with j_rand -> parent-midpoint repair -> fitness -> `<=` selection
// -> reward / cost-curve log -> re-sort + next state features, one CTA per lane, population float32 in
// HBM / float64 in registers.
// Reference: optimizer/lde_optimizer.py:132-140 (init_population), :145-157 (__get_feature),
// :159-198 (update), helpers :21-130.  numpy twin: oracle/lde.py.
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#define LDE_BINS 5

// Shared-memory plan, sized by the run-time population size (NP = 100, D = 30 -> 55.2 KB: 4 CTAs per SM).
// Lifetimes: the sorted parents `ps` (f32) share the evaluation scratch (they are re-read from HBM/L2 in P5); the block
// [sf .. jr] is dead during P3 and doubles as the Gallagher peak-table staging tile.
struct LdeLayout {
    int NP, NPp;
    size_t oConst, oScr, oTs, oF, oTile, oTileEnd, oRank, oSrc, oH, oRed, bytes;
    __host__ __device__ LdeLayout(int np, int D, int kConst, int scratchDoubles) {
        NP = np; NPp = (np + 3) & ~3;
        size_t scr = (size_t)scratchDoubles;
        const size_t psd = ((size_t)NPp * D * sizeof(float) + 7) / 8;
        if (psd > scr) scr = psd;
        if (scr < MBX_NT) scr = MBX_NT;                      // >= 128 x 8 bytes: exchange buffer of the sorting network (P5)
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oScr = oConst + sizeof(double) * kConst;             // P0-P2: parents f32 [NP*D] (sorted)  /  P3: eval scratch f64
        oTs = oScr + sizeof(double) * scr;                   // trial population f32 [NP*D]
        oF = oTs + sizeof(float) * NPp * D;                  // fit, tfit, nfit f32 [NP]
        oTile = (oF + 12 * NPp + 15) & ~(size_t)15;          // sf, cr f32 [NP] | pidx, r1, r2, jr i32 [NP]   (dead in P3)
        oTileEnd = oTile + 24 * NPp;
        oRank = oTileEnd;                                    // rank i32 [NP]
        oSrc = oRank + 4 * NPp;                              // stored row of sorted parent i, u8 [NP]
        oH = (oSrc + NPp + 15) & ~(size_t)15;                // 2 histograms (int[8] each)
        oRed = oH + 64;
        bytes = oRed + 64;
    }
};

template <int D, int KIND>
struct LdeViews {
    MbxInstance* h;
    double *sC, *scr;
    float *ps, *ts, *fit, *tfit, *nfit, *sf, *cr;
    int *pidx, *r1, *r2, *jr, *rank, *hist0, *hist1;
    unsigned char *src, *tile;
    int tile_bytes;
    float* redv;
    int* redi;
    __device__ __forceinline__ LdeViews(unsigned char* s, const LdeLayout& L) {
        const int N = L.NPp;
        h = reinterpret_cast<MbxInstance*>(s);
        sC = reinterpret_cast<double*>(s + L.oConst);
        scr = reinterpret_cast<double*>(s + L.oScr);
        ps = reinterpret_cast<float*>(s + L.oScr);
        ts = reinterpret_cast<float*>(s + L.oTs);
        float* f = reinterpret_cast<float*>(s + L.oF);
        fit = f; tfit = f + N; nfit = f + 2 * N;
        float* g = reinterpret_cast<float*>(s + L.oTile);
        sf = g; cr = g + N;
        int* ii = reinterpret_cast<int*>(s + L.oTile + 8 * N);
        pidx = ii; r1 = ii + N; r2 = ii + 2 * N; jr = ii + 3 * N;
        tile = s + L.oTile; tile_bytes = (int)(L.oTileEnd - L.oTile);
        rank = reinterpret_cast<int*>(s + L.oRank);
        src = s + L.oSrc;
        hist0 = reinterpret_cast<int*>(s + L.oH); hist1 = hist0 + 8;
        redv = reinterpret_cast<float*>(s + L.oRed);
        redi = reinterpret_cast<int*>(s + L.oRed + 32);
    }
};
template <int D, int KIND>
__host__ __device__ inline LdeLayout lde_layout(int NP) {
    return LdeLayout(NP, D, EvalTraits<D, KIND>::kConstDoubles, EvalTraits<D, KIND>::scratch_for((NP + 3) & ~3));
}

// cost-curve protocol (lde_optimizer.py:190-197) -- same tail as every optimizer
__device__ __forceinline__ bool lde_log_and_done(const MbxRunConfig& cfg, bool done, int fes, float gbest, int& log_index,
                                                 int& clen, float* curve) {
    if (fes >= log_index * cfg.log_interval) {
        log_index += 1;
        if (clen < MBX_CURVE_LEN) curve[clen] = gbest;
        clen = min(clen + 1, MBX_CURVE_LEN);
    }
    if (done) {
        if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = gbest;
        else { curve[clen] = gbest; clen += 1; }
    }
    return done;
}

// total order used for the fitness sort: ascending, NaN last, ties by index (stable)
__device__ __forceinline__ bool lde_before(float a, int ia, float b, int ib) {
    const bool na = isnan(a), nb = isnan(b);
    if (na || nb) return (!na && nb) || (na && nb && ia < ib);
    return a < b || (a == b && ia < ib);
}
// rank[i] = position of element i in the stable ascending order of f[0..NP)
__device__ __forceinline__ void lde_rank(const float* f, int* rank, int NP) {
    const int i = threadIdx.x;
    if (i < NP) {
        const float fi = f[i];
        int r = 0;
        for (int j = 0; j < NP; ++j) r += lde_before(f[j], j, fi, i) ? 1 : 0;
        rank[i] = r;
    }
}
// np.histogram(fitness, 5) bin of one value (numpy/lib/_histograms_impl.py fast path; edges = linspace)
__device__ __forceinline__ int lde_bin(double x, double first, double last) {
    const double denom = last - first;
    int idx = (int)(((x - first) / denom) * (double)LDE_BINS);
    if (idx == LDE_BINS) idx -= 1;
    auto edge = [&](int k) { return k == LDE_BINS ? last : (double)k * (denom / (double)LDE_BINS) + first; };
    if (x < edge(idx)) idx -= 1;
    if (x >= edge(idx + 1) && idx != LDE_BINS - 1) idx += 1;
    return idx;
}
// normalised fitness of a float32 cost (lde_optimizer.py:80-86) given the population min / max
__device__ __forceinline__ double lde_norm(float f, float fmin, float fmax) {
    if (fmax != fmin) return ((double)f - (double)fmin) / ((double)fmax - (double)fmin);
    return 0.;
}
// histogram of the normalised fitness of f[0..NP) into hist[0..5); all threads call; includes barriers
__device__ __forceinline__ void lde_hist(const float* f, int NP, float fmin, float fmax, int* hist) {
    if (threadIdx.x < LDE_BINS) hist[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < NP) {
        const bool flat = !(fmax != fmin);
        // np.histogram range = (min, max) of the normalised data = (0,1); flat data (all zeros) widens to (-0.5, 0.5)
        const double first = flat ? -0.5 : 0., last = flat ? 0.5 : 1.;
        atomicAdd(&hist[lde_bin(lde_norm(f[threadIdx.x], fmin, fmax), first, last)], 1);
    }
    __syncthreads();
}
__device__ __forceinline__ void lde_minmax(const float* f, int NP, float* red, float& fmin, float& fmax) {
    float v = threadIdx.x < NP ? f[threadIdx.x] : -INFINITY;
    fmax = block_max<MBX_NT>(v, red);
    v = threadIdx.x < NP ? -f[threadIdx.x] : -INFINITY;
    fmin = -block_max<MBX_NT>(v, red);
}
// features (lde_optimizer.py:145-157): sorted normalised fitness | histogram | mean past histogram
__device__ __forceinline__ void lde_features(const float* fsorted, int NP, float fmin, float fmax, const int* hist,
                                             const double* hsum, int hcnt, float* feat) {
    if (threadIdx.x < NP) feat[threadIdx.x] = (float)lde_norm(fsorted[threadIdx.x], fmin, fmax);
    if (threadIdx.x < LDE_BINS) {
        feat[NP + threadIdx.x] = (float)hist[threadIdx.x];
        feat[NP + LDE_BINS + threadIdx.x] = (float)(hsum[threadIdx.x] / (double)hcnt);
    }
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
lde_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                MbxRunConfig cfg, MbxLdeState S, MbxLdeTape T, float* __restrict__ feature) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, NP = cfg.NP, tid = threadIdx.x;
    const LdeLayout L = lde_layout<D, KIND>(NP);
    LdeViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);
    __syncthreads();
    const MbxInstance& h = *V.h;
    Philox ph(cfg.seed);
    const double lb = h.lb, ub = h.ub;
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    for (int e = tid; e < NP * D; e += MBX_NT) {
        double u = cfg.use_tape ? T.init_u[base + e] : u01(ph((uint32_t)e, MBX_GEN_INIT, (uint32_t)(cfg.lane_base + b), MBX_TAG_INIT).x);
        V.ts[e] = (float)(lb + u * (ub - lb));                             // :67
    }
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.ts, NP, 0, V.scr, V.nfit, true, ph, MBX_GEN_INIT,
                                         (uint32_t)(cfg.lane_base + b), MBX_TAG_NOISE,
                                         cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP, nullptr, 0,
                                         V.tile, V.tile_bytes);
    __syncthreads();
    float fmin, fmax;
    lde_minmax(V.nfit, NP, V.redv, fmin, fmax);
    lde_rank(V.nfit, V.rank, NP);
    __syncthreads();
    if (tid < NP) { V.fit[V.rank[tid]] = V.nfit[tid]; }
    for (int e = tid; e < NP * D; e += MBX_NT) { const int i = e / D; S.pop[base + (size_t)V.rank[i] * D + (e - i * D)] = V.ts[e]; }
    __syncthreads();
    if (tid < NP) S.fit[pb + tid] = V.fit[tid];
    lde_hist(V.fit, NP, fmin, fmax, V.hist1);
    double* hs = S.hist_sum + (size_t)b * LDE_BINS;
    if (tid < LDE_BINS) hs[tid] = (double)NP / (double)LDE_BINS;           // :139
    __syncthreads();
    lde_features(V.fit, NP, fmin, fmax, V.hist1, hs, 1, feature + (size_t)b * (NP + 2 * LDE_BINS));
    if (tid == 0) {
        S.hist_cnt[b] = 1; S.gbest[b] = fmin; S.fes[b] = NP; S.log_index[b] = 1;
        S.curve[(size_t)b * MBX_CURVE_LEN] = fmin; S.curve_len[b] = 1; S.gen[b] = 0; S.done[b] = 0;
    }
}

// =====================================================================================================
// TEST = true: the instantiation the parity tests drive (tape of the oracle's uniforms, debug buffer); TEST = false: what the sweeps launch,
// without the tape's loads / address arithmetic / branches (see rlepso_step_kernel)
template <int D, int KIND, bool TEST>
#ifndef MBX_LDE_MIN_CTAS
#define MBX_LDE_MIN_CTAS 4        // compile-time knob: minimum CTAs per SM of the step kernel (A/B builds: python -m metabox_b200.build --variant)
#endif
__global__ void __launch_bounds__(MBX_NT, MBX_LDE_MIN_CTAS)
lde_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                MbxRunConfig cfg, MbxLdeState S, MbxLdeTape T, const float* __restrict__ action, float* __restrict__ feature,
                float* __restrict__ reward, float* __restrict__ dbg) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }      // finished lane: state frozen, reward 0
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NP = cfg.NP, tid = threadIdx.x;
    const LdeLayout L = lde_layout<D, KIND>(NP);
    LdeViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    const bool use_tape = TEST && cfg.use_tape;
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    const uint32_t gen = (uint32_t)S.gen[b], lane = (uint32_t)(cfg.lane_base + b);
    const int fes0 = S.fes[b];
    Philox ph(cfg.seed);
    // the population rows are requested now (NP * D / 128 <= D values per thread) and scattered into rank order after the rank barrier:
    // their DRAM latency hides behind the constant staging and the rank computation
    float popr[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int e = tid + k * MBX_NT;
        popr[k] = e < NP * D ? __ldcs(S.pop + base + e) : 0.f;
    }
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);          // (after the row requests: its stores wait for its own loads)
    // ---- P0: order_by_f (:160) -- the stored state is sorted already, re-ranked for safety -------------------
    {
        // the stored state IS sorted (P5 below and the init kernel write it in rank order): the rank is the identity unless some pair of
        // neighbours is out of order (a state written by somebody else), and only then is it computed
        const float fi = tid < NP ? S.fit[pb + tid] : 0.f;
        if (tid < NP) V.nfit[tid] = fi;
        __syncthreads();
        const bool bad = tid > 0 && tid < NP && lde_before(fi, tid, V.nfit[tid - 1], tid - 1);
        if (__syncthreads_or(bad ? 1 : 0)) lde_rank(V.nfit, V.rank, NP);
        else if (tid < NP) V.rank[tid] = tid;
    }
    __syncthreads();
    if (tid < NP) { V.fit[V.rank[tid]] = V.nfit[tid]; V.src[V.rank[tid]] = (unsigned char)tid; }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int e = tid + k * MBX_NT;
        if (e < NP * D) { const int i = e / D; V.ps[V.rank[i] * D + (e - i * D)] = popr[k]; }
    }
    // ---- P1: per-individual parameters and draws (:164-167, :88-123) ---------------------------------------------
    if (tid < NP) {
        const int i = tid;
        V.sf[i] = action[(size_t)b * 2 * NP + i];
        V.cr[i] = action[(size_t)b * 2 * NP + NP + i];
        if (use_tape) {
            V.pidx[i] = T.pbest_idx[pb + i]; V.r1[i] = T.r1[pb + i]; V.r2[i] = T.r2[pb + i]; V.jr[i] = T.jrand[pb + i];
        } else {
            const double pmin = 2.0 / (double)NP;                          // P_NUM_MIN / NP
            double p = ((pmin - 1.0) * (double)fes0) / (double)cfg.maxFEs + 1.0;
            p = fmax(0., p);
            const int high = (int)ceil((double)NP * p);
            uint4 w = ph((uint32_t)i, gen, lane, MBX_TAG_LDE_P);
            V.pidx[i] = bounded(w.x, high);
            const int a1 = (i + 1 + bounded(w.y, NP - 1)) % NP;            // uniform over j != i
            int k = bounded(w.z, NP - 2);                                  // uniform over j not in {i, r1}
            const int lo = min(i, a1), hi = max(i, a1);
            k += (k >= lo) ? 1 : 0;
            k += (k >= hi) ? 1 : 0;
            V.r1[i] = a1; V.r2[i] = k; V.jr[i] = bounded(w.w, D);
        }
    }
    __syncthreads();
    const MbxInstance& h = *V.h;
    const double lb = h.lb, ub = h.ub;
    float fmin, fmax;
    lde_minmax(V.fit, NP, V.redv, fmin, fmax);
    lde_hist(V.fit, NP, fmin, fmax, V.hist0);                              // pre-update histogram (:183-187)
    // ---- P2: mutation + crossover + repair, element-wise over (i,d) (:44-51, :88-130) ------------------------------
    for (int e = tid; e < NP * D; e += MBX_NT) {
        const int i = e / D, d = e - i * D;
        const float sff = V.sf[i];
        const double xi = (double)V.ps[e], sf = (double)sff;
        const int p = V.pidx[i];
        double mu = xi;
        if (p != i) mu = (double)(1.f - sff) * xi + sf * (double)V.ps[p * D + d];
        mu = mu + sf * ((double)V.ps[V.r1[i] * D + d] - (double)V.ps[V.r2[i] * D + d]);
        double u = use_tape ? T.cross_u[base + e] : u01(ph((uint32_t)e, gen, lane, MBX_TAG_LDE_PD).x);
        if (d == V.jr[i]) u = 0.;
        double t = (u <= (double)V.cr[i]) ? mu : xi;
        if (t < lb) t = (xi + lb) / 2.;
        else if (t > ub) t = (xi + ub) / 2.;
        const float t32 = (float)t;
        V.ts[e] = t32;
        if (TEST && dbg) dbg[(size_t)b * NP * (D + 1) + e] = t32;
    }
    __syncthreads();
    // ---- P3: fitness of the trials (:54) -------------------------------------------------------------------------------
    // (the evaluation scratch overwrites the parents' shared-memory copy; P5 re-reads the surviving parents from HBM/L2)
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.ts, NP, 0, V.scr, V.tfit, true, ph, gen, lane,
                                         MBX_TAG_NOISE, use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP,
                                         nullptr, 0, V.tile, V.tile_bytes);
    __syncthreads();
    // ---- P4: selection `<=` (:59-61), reward (:170-173) ------------------------------------------------------------------
    if (tid < NP) {
        const float tf = V.tfit[tid];
        if (TEST && dbg) dbg[(size_t)b * NP * (D + 1) + NP * D + tid] = tf;
        V.nfit[tid] = (tf <= V.fit[tid]) ? tf : V.fit[tid];
    }
    __syncthreads();
    float nmin, nmax;
    lde_minmax(V.nfit, NP, V.redv, nmin, nmax);
    const int fes = fes0 + NP;
    const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)nmin <= 1e-8) : (fes >= cfg.maxFEs);
    // ---- P5: re-sort the survivors (__get_feature :146) and write the state back --------------------------------------------
    {
        // (the evaluation scratch is dead: exchange buffer of the sorting network)
        const int at = block_sort_order_128(tid < NP ? V.nfit[tid] : 0.f, NP, reinterpret_cast<unsigned long long*>(V.scr));
        if (tid < NP) V.rank[at] = tid;
    }
    __syncthreads();
    {
        // survivors: trial if tfit <= fit else the parent, whose row is re-read from the stored (old) population; every
        // read of the old population happens before the first write of the new one
        constexpr int EPT = (MBX_MAX_NP * D + MBX_NT - 1) / MBX_NT;
        float keep[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int e = tid + k * MBX_NT;
            keep[k] = 0.f;
            if (e < NP * D) {
                const int i = e / D;
                keep[k] = (V.tfit[i] <= V.fit[i]) ? V.ts[e] : S.pop[base + (size_t)V.src[i] * D + (e - i * D)];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int e = tid + k * MBX_NT;
            if (e < NP * D) { const int i = e / D; S.pop[base + (size_t)V.rank[i] * D + (e - i * D)] = keep[k]; }
        }
    }
    __syncthreads();
    if (tid < NP) V.tfit[V.rank[tid]] = V.nfit[tid];                       // tfit := survivors' costs, sorted
    __syncthreads();
    if (tid < NP) S.fit[pb + tid] = V.tfit[tid];
    lde_hist(V.tfit, NP, nmin, nmax, V.hist1);
    double* hs = S.hist_sum + (size_t)b * LDE_BINS;
    const int hcnt = S.hist_cnt[b] + 1;
    if (tid < LDE_BINS) hs[tid] = hs[tid] + (double)V.hist0[tid];
    __syncthreads();
    lde_features(V.tfit, NP, nmin, nmax, V.hist1, hs, hcnt, feature + (size_t)b * (NP + 2 * LDE_BINS));
    if (tid == 0) {
        int lg = S.log_index[b], clen = S.curve_len[b];
        lde_log_and_done(cfg, done, fes, nmin, lg, clen, S.curve + (size_t)b * MBX_CURVE_LEN);
        S.log_index[b] = lg; S.curve_len[b] = clen; S.fes[b] = fes; S.gen[b] = (int)gen + 1; S.hist_cnt[b] = hcnt;
        S.gbest[b] = nmin; S.done[b] = done ? 1 : 0;
        reward[b] = (float)(((double)fmin - (double)nmin) / (double)fmin);
    }
}

// =====================================================================================================
template <int D, int KIND>
static int lde_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                             const MbxLdeState* s, const MbxLdeTape* tape, float* feature, int B, cudaStream_t st) {
    auto kern = lde_init_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done.set();
    }
    MbxLdeTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, lde_layout<D, KIND>(cfg->NP).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, feature);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND, bool TEST>
static int lde_launch_step_tt(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                              const MbxLdeState* s, const MbxLdeTape* tape, const float* action, float* feature, float* reward,
                              float* dbg, int B, cudaStream_t st) {
    auto kern = lde_step_kernel<D, KIND, TEST>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done.set();
    }
    MbxLdeTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, lde_layout<D, KIND>(cfg->NP).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, feature, reward, dbg);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int lde_launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                             const MbxLdeState* s, const MbxLdeTape* tape, const float* action, float* feature, float* reward,
                             float* dbg, int B, cudaStream_t st) {
    if (cfg->use_tape || dbg) return lde_launch_step_tt<D, KIND, true>(inst, consts, lane_inst, cfg, s, tape, action, feature, reward, dbg, B, st);
    return lde_launch_step_tt<D, KIND, false>(inst, consts, lane_inst, cfg, s, tape, action, feature, reward, dbg, B, st);
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_lde_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                        const MbxRunConfig* cfg, const MbxLdeState* s, const MbxLdeTape* tape, float* feature, int B,
                        cudaStream_t st) {
    MBX_DISPATCH(lde_launch_init_t, inst, consts, lane_inst, cfg, s, tape, feature, B, st)
    mbx_set_error("mbx_lde_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_lde_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                        const MbxRunConfig* cfg, const MbxLdeState* s, const MbxLdeTape* tape, const float* action,
                        float* feature, float* reward, float* dbg, int B, cudaStream_t st) {
    MBX_DISPATCH(lde_launch_step_t, inst, consts, lane_inst, cfg, s, tape, action, feature, reward, dbg, B, st)
    mbx_set_error("mbx_lde_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
