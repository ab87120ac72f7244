// mbx_gleet.cu -- fused GLEET generation (PSO family, whole swarm per step): inertia decay -> velocity from the
// per-particle exploration/exploitation action -> clip -> position -> fitness -> pbest/gbest -> patience counters ->
// the 9 observation features per particle (costs, budget, stagnation, distances to gbest / pbest, cosine of the two
// attraction directions) -> carried pbest / gbest features -> [NP, 27] state, reward, cost-curve log.  One CTA per
// lane, state float32 in HBM / float64 in registers.
// Reference: optimizer/gleet_optimizer.py:34-60 (initialize_particles), :76-112 (init_population), :128-158 (observe,
// gp_cat), :193-314 (update); fixed configuration ps = 100, c = 4.1, w = 0.9 decaying, clipping border, 'direct'
// reward x 100 (:10-32).  numpy twin: oracle/gleet.py.
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#define GL_NFEAT 9
#define GL_C 4.1f
#define GL_REWARD_SCALE 100.

// Shared-memory plan (NP = 100, D = 30 -> 55.9 KB: 4 CTAs per SM).  Region A holds the pbest positions (f32) while the
// velocities are formed (P1) and again for the features (P4), and is the evaluation scratch in between (P2).
struct GlLayout {
    int NP, NPp;
    size_t oConst, oA, oXs, oTile, oTileEnd, oF, oPni, oGx, oGf, oRed, bytes;
    __host__ __device__ GlLayout(int np, int D, int kConst, int scratchDoubles) {
        NP = np; NPp = (np + 3) & ~3;
        size_t scr = (size_t)scratchDoubles;
        const size_t pbd = ((size_t)NPp * (D + 27) * sizeof(float) + 7) / 8;   // pbest_x + the [NP, 27] state staging of P4
        if (pbd > scr) scr = pbd;
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oA = oConst + sizeof(double) * kConst;               // pbest_x f32 [NP*D]  /  eval scratch f64
        oXs = oA + sizeof(double) * scr;                     // new positions f32 [NP*D]
        oTile = oXs + sizeof(float) * NPp * D;               // r1, r2 (Philox words u32), ca, cb (f32) [NP]  (dead in P2)
        oTileEnd = oTile + 16 * NPp;
        oF = oTileEnd;                                       // pf32, nc, cc f32 [NP]
        oPni = oF + 12 * NPp;                                // per_no_improve i32 [NP] | better u8 [NP]
        oGx = (oPni + 5 * NPp + 7) & ~(size_t)7;             // gbest_x f64 [D]
        oGf = oGx + sizeof(double) * D;                      // gbest feature f32 [9] (+ pad)
        oRed = (oGf + 48 + 15) & ~(size_t)15;                // reduction scratch, w
        bytes = oRed + 96;
    }
};
template <int D, int KIND>
__host__ __device__ inline GlLayout gl_layout(int NP) {
    return GlLayout(NP, D, EvalTraits<D, KIND>::kConstDoubles, EvalTraits<D, KIND>::scratch_for((NP + 3) & ~3));
}
template <int D, int KIND>
struct GlViews {
    MbxInstance* h;
    double *sC, *A, *gx, *sw;
    float *pbx, *xs, *ca, *cb, *pf32, *nc, *cc, *gf;
    uint32_t *r1u, *r2u;
    int* pni;
    unsigned char *better, *tile;
    int tile_bytes;
    float* redv;
    int* redi;
    __device__ __forceinline__ GlViews(unsigned char* s, const GlLayout& L) {
        const int N = L.NPp;
        h = reinterpret_cast<MbxInstance*>(s);
        sC = reinterpret_cast<double*>(s + L.oConst);
        A = reinterpret_cast<double*>(s + L.oA);
        pbx = reinterpret_cast<float*>(s + L.oA);
        xs = reinterpret_cast<float*>(s + L.oXs);
        r1u = reinterpret_cast<uint32_t*>(s + L.oTile); r2u = r1u + N;
        ca = reinterpret_cast<float*>(s + L.oTile + 8 * N); cb = ca + N;
        tile = s + L.oTile; tile_bytes = (int)(L.oTileEnd - L.oTile);
        float* f = reinterpret_cast<float*>(s + L.oF);
        pf32 = f; nc = f + N; cc = f + 2 * N;
        pni = reinterpret_cast<int*>(s + L.oPni);
        better = s + L.oPni + 4 * N;
        gx = reinterpret_cast<double*>(s + L.oGx);
        gf = reinterpret_cast<float*>(s + L.oGf);
        redv = reinterpret_cast<float*>(s + L.oRed);
        redi = reinterpret_cast<int*>(s + L.oRed + 32);
        sw = reinterpret_cast<double*>(s + L.oRed + 64);
    }
};

// observe() for particle i (gleet_optimizer.py:128-154): 9 float64 features from the float32 state in shared memory
template <int D>
__device__ __forceinline__ void gl_observe(const MbxRunConfig& cfg, const float* xs, const float* pbx, const double* gx, const int i,
                                           const double cc, const double pb, const double gb, const double mc, const int pni,
                                           const int no_improve, const int fes, const double lb, const double ub, double (&f)[GL_NFEAT]) {
    const double max_step = (double)(cfg.maxFEs / cfg.NP);                        // integer division (:130)
    const double max_dist = sqrt((ub - lb) * (ub - lb) * (double)D);               // :84
    f[0] = cc / mc;
    f[1] = (cc - gb) / mc;
    f[2] = (cc - pb) / mc;
    f[3] = (double)(cfg.maxFEs - fes) / (double)cfg.maxFEs;
    f[4] = (double)pni / max_step;
    f[5] = (double)no_improve / max_step;
    double dg = 0., dp = 0., dot = 0.;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
        const double x = (double)xs[i * D + d];
        const double pv = (double)pbx[i * D + d] - x, gv = gx[d] - x;
        dg += gv * gv; dp += pv * pv; dot += pv * gv;
    }
    f[6] = sqrt(dg) / max_dist;
    f[7] = sqrt(dp) / max_dist;
    const double c8 = dot / (sqrt(dp) * sqrt(dg) + 1e-5);
    f[8] = (c8 != c8) ? 0. : c8;
}

// cost-curve protocol shared by all optimizers (gleet_optimizer.py:303-312)
__device__ __forceinline__ void gl_log(const MbxRunConfig& cfg, bool done, int fes, float gbest, int& log_index, int& clen, float* curve) {
    if (fes >= log_index * cfg.log_interval) {
        log_index += 1;
        if (clen < MBX_CURVE_LEN) curve[clen] = gbest;
        clen = min(clen + 1, MBX_CURVE_LEN);
    }
    if (done) {
        if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = gbest;
        else { curve[clen] = gbest; clen += 1; }
    }
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
gleet_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxGleetState S, MbxGleetTape T, float* __restrict__ state) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, NP = cfg.NP, tid = threadIdx.x;
    const GlLayout L = gl_layout<D, KIND>(NP);
    GlViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);
    __syncthreads();
    const MbxInstance& h = *V.h;
    Philox ph(cfg.seed);
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb);
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    for (int e = tid; e < NP * D; e += MBX_NT) {
        double up, uv;
        if (cfg.use_tape) { up = T.init_pos_u[base + e]; uv = T.init_vel_u[base + e]; }
        else { uint4 w = ph((uint32_t)e, MBX_GEN_INIT, lane, MBX_TAG_INIT); up = u01(w.x); uv = u01(w.y); }
        const float p = (float)(lb + (ub - lb) * up);
        const float v = (float)(-vmax + (vmax - (-vmax)) * uv);
        V.xs[e] = p;
        S.x[base + e] = p; S.pbest_x[base + e] = p; S.v[base + e] = v;
    }
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, V.nc, true, ph, MBX_GEN_INIT, lane, MBX_TAG_NOISE,
                                         cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP, nullptr, 0,
                                         V.tile, V.tile_bytes);
    __syncthreads();
    float v = INFINITY; int ix = tid;
    if (tid < NP) {
        v = V.nc[tid];
        S.c_cost[pb + tid] = v; S.pbest[pb + tid] = v; S.pni[pb + tid] = 0;
    }
    block_argmin<MBX_NT>(v, ix, V.redv, V.redi);
    for (int e = tid; e < NP * D; e += MBX_NT) V.pbx[e] = V.xs[e];          // pbest = current (scratch is free again)
    for (int d = tid; d < D; d += MBX_NT) { V.gx[d] = (double)V.xs[ix * D + d]; S.gbest_x[(size_t)b * D + d] = V.xs[ix * D + d]; }
    __syncthreads();
    float* out = state + ((size_t)b * NP) * 3 * GL_NFEAT;
    float myf[GL_NFEAT];
    if (tid < NP) {
        double f[GL_NFEAT];
        const double c = (double)V.nc[tid];
        gl_observe<D>(cfg, V.xs, V.pbx, V.gx, tid, c, c, (double)v, (double)v, 0, 0, NP, lb, ub, f);
#pragma unroll
        for (int k = 0; k < GL_NFEAT; ++k) {
            myf[k] = (float)f[k];
            out[tid * 3 * GL_NFEAT + k] = myf[k]; out[tid * 3 * GL_NFEAT + GL_NFEAT + k] = myf[k];
            S.pbest_feature[(pb + tid) * GL_NFEAT + k] = myf[k];
        }
        if (tid == ix) {
#pragma unroll
            for (int k = 0; k < GL_NFEAT; ++k) { V.gf[k] = myf[k]; S.gbest_feature[(size_t)b * GL_NFEAT + k] = myf[k]; }
        }
    }
    __syncthreads();
    if (tid < NP) {
#pragma unroll
        for (int k = 0; k < GL_NFEAT; ++k) out[tid * 3 * GL_NFEAT + 2 * GL_NFEAT + k] = V.gf[k];
    }
    if (tid == 0) {
        S.gbest_val[b] = v; S.gbest_idx[b] = ix; S.no_improve[b] = 0; S.fes[b] = NP; S.log_index[b] = 1; S.w[b] = 0.9;
        S.max_cost[b] = v;                                                   // `self.max_cost = np.min(c_cost)` (:49)
        S.curve[(size_t)b * MBX_CURVE_LEN] = v; S.curve_len[b] = 1; S.gen[b] = 0; S.done[b] = 0;
    }
}

// =====================================================================================================
// TEST = true: the instantiation the parity tests drive (tape of the oracle's uniforms); TEST = false: what the sweeps launch, without the
// tape's loads / address arithmetic / branches (see rlepso_step_kernel)
template <int D, int KIND, bool TEST>
#ifndef MBX_GLEET_MIN_CTAS
#define MBX_GLEET_MIN_CTAS 4        // compile-time knob: minimum CTAs per SM of the step kernel (A/B builds: python -m metabox_b200.build --variant)
#endif
__global__ void __launch_bounds__(MBX_NT, MBX_GLEET_MIN_CTAS)
gleet_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxGleetState S, MbxGleetTape T, const float* __restrict__ action, float* __restrict__ state,
                  float* __restrict__ reward) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }     // finished lane: state frozen, reward 0
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NP = cfg.NP, tid = threadIdx.x;
    const GlLayout L = gl_layout<D, KIND>(NP);
    GlViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    const bool use_tape = TEST && cfg.use_tape;
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    // the pbest rows are requested first (NP * D / 128 <= D values per thread): their DRAM latency hides behind the constant staging and P0
    float pr[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { const int e = tid + k * MBX_NT; pr[k] = e < NP * D ? __ldcs(S.pbest_x + base + e) : 0.f; }
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);
    const uint32_t gen = (uint32_t)S.gen[b], lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    // ---- P0: scalars, coefficients (float32 like numpy-2's promotion of `self.c * action`), draws, pbest tile ----------------
    for (int d = tid; d < D; d += MBX_NT) V.gx[d] = (double)S.gbest_x[(size_t)b * D + d];
    if (tid == 0) {
        const double w = S.w[b] - 0.5 / ((double)cfg.maxFEs / (double)NP);   // :201-202, the same float64 chain
        *V.sw = w; S.w[b] = w;
    }
    if (tid < NP) {
        const int i = tid;
        V.pf32[i] = S.pbest[pb + i]; V.cc[i] = S.c_cost[pb + i]; V.pni[i] = S.pni[pb + i];
        const float a = action[pb + i];
        V.ca[i] = GL_C * a;
        V.cb[i] = GL_C * (1.f - a);
        if (!use_tape) { uint4 w4 = ph((uint32_t)i, gen, lane, MBX_TAG_GLEET_P); V.r1u[i] = w4.x; V.r2u[i] = w4.y; }
    }
    {
#pragma unroll
        for (int k = 0; k < D; ++k) { const int e = tid + k * MBX_NT; if (e < NP * D) V.pbx[e] = pr[k]; }
    }
    __syncthreads();
    const MbxInstance& h = *V.h;
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb), w = *V.sw;
    // ---- P1: velocity + position, element-wise over (i,d)  (:204-221) -----------------------------------------------------
    constexpr int CH = 12;
#pragma unroll 1
    for (int k0 = 0; k0 * MBX_NT < NP * D; k0 += CH) {
        float xr[CH], vr[CH];
        // one 64-bit pointer per array and chunk; the elements of the chunk sit at compile-time offsets from it
        float* const xg = S.x + base + tid + k0 * MBX_NT;
        float* const vg = S.v + base + tid + k0 * MBX_NT;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int e = tid + (k0 + k) * MBX_NT;
            xr[k] = 0.f; vr[k] = 0.f;
            if (e < NP * D) { xr[k] = __ldcs(xg + k * MBX_NT); vr[k] = __ldcs(vg + k * MBX_NT); }
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int e = tid + (k0 + k) * MBX_NT;
            if (e >= NP * D) break;
            const int i = e / D, d = e - i * D;
            const double x = (double)xr[k], v = (double)vr[k];
            const double r1 = use_tape ? T.rand1[pb + i] : u01(V.r1u[i]), r2 = use_tape ? T.rand2[pb + i] : u01(V.r2u[i]);
            double nv = w * v + (double)V.ca[i] * r1 * ((double)V.pbx[e] - x) + (double)V.cb[i] * r2 * (V.gx[d] - x);
            // np.clip as two compares and selects each (fmin / fmax add NaN handling the reference does not have: np.clip keeps a NaN)
            nv = nv < -vmax ? -vmax : nv; nv = nv > vmax ? vmax : nv;
            double nx = x + nv;
            nx = nx < lb ? lb : nx; nx = nx > ub ? ub : nx;
            const float nx32 = (float)nx;
            V.xs[e] = nx32;
            xg[k * MBX_NT] = nx32; vg[k * MBX_NT] = (float)nv;
        }
    }
    __syncthreads();
    // ---- P2: fitness (:224) ------------------------------------------------------------------------------------------------------
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, V.nc, true, ph, gen, lane, MBX_TAG_NOISE,
                                         use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP, nullptr, 0,
                                         V.tile, V.tile_bytes);
    __syncthreads();
    // ---- P3: pbest / gbest / patience (:227-274) ------------------------------------------------------------------------------
    const int fes = S.fes[b] + NP;
    const float pre_gbest = S.gbest_val[b];
    float gbest = pre_gbest;
    float v = INFINITY; int ix = tid;
    // the particle's carried exploration features (read in P4 unless it improved) are requested now: one DRAM latency behind the whole of P3
    float pfe[GL_NFEAT];
#pragma unroll
    for (int k = 0; k < GL_NFEAT; ++k) pfe[k] = tid < NP ? S.pbest_feature[(pb + tid) * GL_NFEAT + k] : 0.f;
    if (tid < NP) {
        const int i = tid;
        const float ncv = V.nc[i];
        v = ncv;
        const bool bt = ncv < V.pf32[i];
        V.better[i] = bt;
        if (bt) { V.pf32[i] = ncv; S.pbest[pb + i] = ncv; }
        S.c_cost[pb + i] = ncv;
        const int pni = (ncv < V.cc[i]) ? 0 : V.pni[i] + 1;
        V.pni[i] = pni; S.pni[pb + i] = pni;
        V.cc[i] = ncv;
    }
    block_argmin<MBX_NT>(v, ix, V.redv, V.redi);                  // (contains barriers: V.better is visible below)
    int gidx = S.gbest_idx[b];
    const bool gb_better = v < gbest;
    if (gb_better) { gbest = v; gidx = ix; }
    const int no_improve = gb_better ? 0 : S.no_improve[b] + 1;
    // the evaluation scratch is free again: rebuild the pbest tile (old rows from HBM/L2, improved rows from xs)
    {   // loads first, stores after: S.pbest_x is read and written through the same pointer, so interleaved iterations would serialise
        float pv[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int e = tid + k * MBX_NT;
            pv[k] = 0.f;
            if (e < NP * D) pv[k] = V.better[e / D] ? V.xs[e] : S.pbest_x[base + e];
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int e = tid + k * MBX_NT;
            if (e < NP * D) {
                if (V.better[e / D]) S.pbest_x[base + e] = pv[k];
                V.pbx[e] = pv[k];
            }
        }
    }
    if (gb_better)
        for (int d = tid; d < D; d += MBX_NT) { V.gx[d] = (double)V.xs[ix * D + d]; S.gbest_x[(size_t)b * D + d] = V.xs[ix * D + d]; }
    __syncthreads();
    // ---- P4: observe + carried features -> state [NP, 27]  (:128-158, :297-300) --------------------------------------------------
    const double mc = (double)S.max_cost[b];
    float* out = state + pb * 3 * GL_NFEAT;
    // the [NP, 27] state is assembled in shared memory (second half of region A, free here) and written out coalesced
    float* stage = V.pbx + (size_t)L.NPp * D;
    if (tid < NP) {
        const int i = tid;
        double f[GL_NFEAT];
        gl_observe<D>(cfg, V.xs, V.pbx, V.gx, i, (double)V.cc[i], (double)V.pf32[i], (double)gbest, mc, V.pni[i], no_improve, fes, lb, ub, f);
        float* o = stage + i * 3 * GL_NFEAT;
        float* pfeat = S.pbest_feature + (pb + i) * GL_NFEAT;
        const bool fresh = V.pni[i] == 0;                          // exploration state follows the particle while it improves
#pragma unroll
        for (int k = 0; k < GL_NFEAT; ++k) {
            const float fk = (float)f[k];
            o[k] = fk;
            if (fresh) { pfeat[k] = fk; o[GL_NFEAT + k] = fk; }
            else o[GL_NFEAT + k] = pfe[k];
        }
        if (i == gidx) {
#pragma unroll
            for (int k = 0; k < GL_NFEAT; ++k) {
                if (no_improve == 0) { V.gf[k] = (float)f[k]; S.gbest_feature[(size_t)b * GL_NFEAT + k] = (float)f[k]; }
            }
        }
    }
    if (no_improve != 0 && tid < GL_NFEAT) V.gf[tid] = S.gbest_feature[(size_t)b * GL_NFEAT + tid];
    __syncthreads();
    for (int e = tid; e < NP * 3 * GL_NFEAT; e += MBX_NT) {
        const int k = e % (3 * GL_NFEAT);
        out[e] = (k >= 2 * GL_NFEAT) ? V.gf[k - 2 * GL_NFEAT] : stage[e];
    }
    // ---- P5: bookkeeping (:277-312) --------------------------------------------------------------------------------------------
    if (tid == 0) {
        const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)gbest <= 1e-8) : (fes >= cfg.maxFEs);
        if (gb_better) { S.gbest_val[b] = gbest; S.gbest_idx[b] = gidx; }
        S.no_improve[b] = no_improve;
        int lg = S.log_index[b], clen = S.curve_len[b];
        gl_log(cfg, done, fes, gbest, lg, clen, S.curve + (size_t)b * MBX_CURVE_LEN);
        S.log_index[b] = lg; S.curve_len[b] = clen; S.fes[b] = fes; S.gen[b] = (int)gen + 1; S.done[b] = done ? 1 : 0;
        reward[b] = (float)(((double)pre_gbest - (double)gbest) / mc * GL_REWARD_SCALE);      // 'direct' x reward_scale
    }
}

// =====================================================================================================
template <int D, int KIND>
static int gl_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxGleetState* s, const MbxGleetTape* tape, float* state, int B, cudaStream_t st) {
    auto kern = gleet_init_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done.set();
    }
    MbxGleetTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, gl_layout<D, KIND>(cfg->NP).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, state);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND, bool TEST>
static int gl_launch_step_tt(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                             const MbxGleetState* s, const MbxGleetTape* tape, const float* action, float* state, float* reward,
                             int B, cudaStream_t st) {
    auto kern = gleet_step_kernel<D, KIND, TEST>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done.set();
    }
    MbxGleetTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, gl_layout<D, KIND>(cfg->NP).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, state, reward);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int gl_launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxGleetState* s, const MbxGleetTape* tape, const float* action, float* state, float* reward,
                            int B, cudaStream_t st) {
    if (cfg->use_tape) return gl_launch_step_tt<D, KIND, true>(inst, consts, lane_inst, cfg, s, tape, action, state, reward, B, st);
    return gl_launch_step_tt<D, KIND, false>(inst, consts, lane_inst, cfg, s, tape, action, state, reward, B, st);
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_gleet_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxGleetState* s, const MbxGleetTape* tape, float* state, int B,
                          cudaStream_t st) {
    MBX_DISPATCH(gl_launch_init_t, inst, consts, lane_inst, cfg, s, tape, state, B, st)
    mbx_set_error("mbx_gleet_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_gleet_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxGleetState* s, const MbxGleetTape* tape, const float* action,
                          float* state, float* reward, int B, cudaStream_t st) {
    MBX_DISPATCH(gl_launch_step_t, inst, consts, lane_inst, cfg, s, tape, action, state, reward, B, st)
    mbx_set_error("mbx_gleet_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
