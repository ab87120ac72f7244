// mbx_rlpso.cu -- fused RL_PSO step (PSO family, ONE particle per step): inertia decay -> velocity of the current particle
// from (rand1, action) -> clip -> position -> fitness of the single candidate (CTA-cooperative, one dimension per thread)
// -> pbest / gbest -> reward -> cost-curve log -> state of the NEXT particle [gbest_position | current_position].
// One CTA per lane; state float32 in HBM / float64 in registers.
// Reference: optimizer/rl_pso_optimizer.py:32-62 (init_population), :64-65 (__get_state), :79-148 (update); fixed
// configuration NP = 100, c = 2.05, w = 0.9 decaying by 0.5 / (maxFEs / NP) per step (:10-24).  numpy twin: oracle/rl_pso.py.
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#define RP_C 2.05

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
rlpso_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxRlPsoState S, MbxRlPsoTape T, float* __restrict__ state) {
    using L = PopSmem<D, KIND>;                                            // NP candidates at once: the population layout
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw + L::oHdr);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oA);
    float* xs = reinterpret_cast<float*>(smem_raw + L::oXs);
    float* sc = reinterpret_cast<float*>(smem_raw + L::oF);
    float* redv = reinterpret_cast<float*>(smem_raw + L::oRed);
    int* redi = reinterpret_cast<int*>(smem_raw + L::oRed + 32);
    const int b = blockIdx.x, NP = cfg.NP, tid = threadIdx.x;
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, sh, sC);
    __syncthreads();
    const MbxInstance& h = *sh;
    Philox ph(cfg.seed);
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb);
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    for (int e = tid; e < NP * D; e += MBX_NT) {
        double up, uv;
        if (cfg.use_tape) { up = T.init_pos_u[base + e]; uv = T.init_vel_u[base + e]; }
        else { uint4 w = ph((uint32_t)e, MBX_GEN_INIT, lane, MBX_TAG_INIT); up = u01(w.x); uv = u01(w.y); }
        const float p = (float)(lb + (ub - lb) * up);
        xs[e] = p;
        S.x[base + e] = p; S.pbest_x[base + e] = p; S.v[base + e] = (float)(-vmax + (vmax - (-vmax)) * uv);
    }
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, sC, consts + gi->off, xs, NP, 0, scratch, sc, true, ph, MBX_GEN_INIT, lane, MBX_TAG_NOISE,
                                         cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP);
    __syncthreads();
    float v = INFINITY; int ix = tid;
    if (tid < NP) { v = sc[tid]; S.c_cost[pb + tid] = v; S.pbest[pb + tid] = v; }
    const float worst = block_max<MBX_NT>(tid < NP ? v : -INFINITY, redv);
    block_argmin<MBX_NT>(v, ix, redv, redi);
    for (int d = tid; d < D; d += MBX_NT) {
        S.gbest_x[(size_t)b * D + d] = xs[ix * D + d];
        state[(size_t)b * 2 * D + d] = xs[ix * D + d];                     // gbest_position | current_position[0]  (:64-65)
        state[(size_t)b * 2 * D + D + d] = xs[d];
    }
    if (tid == 0) {
        S.gbest_val[b] = v; S.gbest_idx[b] = ix; S.max_cost[b] = worst; S.w[b] = 0.9; S.cur_index[b] = 0;
        S.fes[b] = NP; S.log_index[b] = 1; S.curve[(size_t)b * MBX_CURVE_LEN] = v; S.curve_len[b] = 1; S.done[b] = 0;
    }
}

// =====================================================================================================
template <int D, int KIND>
struct RpSmem {
    static constexpr int kConst = (KIND == MBX_KIND_PROTEIN) ? EvalTraits<D, KIND>::kConstDoubles : 0;   // BBOB constants stay in global
    static constexpr int kScratch = (KIND == MBX_KIND_PROTEIN) ? MBX_PROT_WARP_SCRATCH : D + 1;
    static constexpr size_t oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
    static constexpr size_t oScr = oConst + sizeof(double) * kConst;
    static constexpr size_t oTrial = oScr + sizeof(double) * kScratch;
    static constexpr size_t oMisc = oTrial + sizeof(float) * MBX_MAX_DIM;
    static constexpr size_t kBytes = ((oMisc + 64 + 15) / 16) * 16;
};

template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT, 8)
rlpso_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxRlPsoState S, MbxRlPsoTape T, const float* __restrict__ action, float* __restrict__ state,
                  float* __restrict__ reward) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }      // finished lane: state frozen, reward 0
    using L = RpSmem<D, KIND>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oScr);
    float* trial = reinterpret_cast<float*>(smem_raw + L::oTrial);
    float* mf = reinterpret_cast<float*>(smem_raw + L::oMisc);             // [0] new cost
    int* mi = reinterpret_cast<int*>(smem_raw + L::oMisc + 16);            // [0] pbest improved, [1] gbest improved
    const int NP = cfg.NP, tid = threadIdx.x;
    const MbxInstance* gi = inst + lane_inst[b];
    const double* cC = sC;
    if (KIND == MBX_KIND_PROTEIN) stage_instance<D, KIND>(gi, consts, sh, sC);
    else {                                                                 // one candidate per step: constants straight from global / L1
        const int* hs = reinterpret_cast<const int*>(gi);
        int* hd = reinterpret_cast<int*>(sh);
        for (int k = tid; k < (int)(sizeof(MbxInstance) / 4); k += MBX_NT) hd[k] = hs[k];
        cC = consts + gi->off;
    }
    const int j = S.cur_index[b], fes0 = S.fes[b];
    const size_t row = ((size_t)b * NP + j) * D;
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    const double w = S.w[b] - 0.5 / ((double)cfg.maxFEs / (double)NP);     // :86-87, the same float64 chain
    const float pre_cost = S.c_cost[(size_t)b * NP + j];
    __syncthreads();
    const MbxInstance& h = *sh;
    const double lb = h.lb, ub = h.ub, vmax = 0.1 * (ub - lb);
    // ---- velocity + position of particle j (:90-112); numpy-2 promotion: `self.__c * rand2` is float32 ----------------
    if (tid < D) {
        const int d = tid;
        const double rand1 = cfg.use_tape ? T.rand1[b] : u01(ph(0u, (uint32_t)fes0, lane, MBX_TAG_RLPSO).x);
        const float c2 = (float)RP_C * action[b];
        const double x = (double)S.x[row + d], v = (double)S.v[row + d], pbx = (double)S.pbest_x[row + d];
        const double gx = (double)S.gbest_x[(size_t)b * D + d];
        double nv = w * v + RP_C * rand1 * (pbx - x) + (double)c2 * (gx - x);
        nv = nv < -vmax ? -vmax : nv; nv = nv > vmax ? vmax : nv;      // np.clip (keeps a NaN, unlike fmin / fmax)
        double nx = x + nv;
        nx = nx < lb ? lb : nx; nx = nx > ub ? ub : nx;
        const float nx32 = (float)nx;
        trial[d] = nx32;
        S.x[row + d] = nx32; S.v[row + d] = (float)nv;
    }
    __syncthreads();
    // ---- fitness of the single candidate (:115) -----------------------------------------------------------------------------
    {
        const double f = cta_eval_one<D, KIND>(h, cC, consts + gi->off, trial, scratch, ph, 0u, (uint32_t)fes0, lane,
                                               cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 : nullptr, 1);
        if (tid == 0) {
            const float nc = (float)f;
            mf[0] = nc;
            S.c_cost[(size_t)b * NP + j] = nc;
            const bool bp = nc < S.pbest[(size_t)b * NP + j];
            if (bp) S.pbest[(size_t)b * NP + j] = nc;
            mi[0] = bp ? 1 : 0;
            mi[1] = (nc < S.gbest_val[b]) ? 1 : 0;
        }
    }
    __syncthreads();
    const float nc = mf[0];
    const bool bp = mi[0] != 0, bg = mi[1] != 0;
    const int jn = (j + 1) % NP;
    // ---- pbest / gbest rows (:118-125) and the state of the next particle (:64-65) ------------------------------------------------
    if (tid < D) {
        const int d = tid;
        if (bp) S.pbest_x[row + d] = trial[d];
        float gxn = S.gbest_x[(size_t)b * D + d];
        if (bg) { gxn = trial[d]; S.gbest_x[(size_t)b * D + d] = gxn; }
        state[(size_t)b * 2 * D + d] = gxn;
        state[(size_t)b * 2 * D + D + d] = (jn == j) ? trial[d] : S.x[((size_t)b * NP + jn) * D + d];
    }
    if (tid == 0) {
        float gbest = S.gbest_val[b];
        if (bg) { gbest = nc; S.gbest_val[b] = nc; S.gbest_idx[b] = j; }
        const int fes = fes0 + 1;
        const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)gbest <= 1e-8) : (fes >= cfg.maxFEs);
        int lg = S.log_index[b], clen = S.curve_len[b];
        float* curve = S.curve + (size_t)b * MBX_CURVE_LEN;
        if (fes >= lg * cfg.log_interval) {                                // :133-135
            lg += 1;
            if (clen < MBX_CURVE_LEN) curve[clen] = gbest;
            clen = min(clen + 1, MBX_CURVE_LEN);
        }
        reward[b] = (float)(((double)pre_cost - (double)nc) / ((double)S.max_cost[b] - (double)gbest));   // :137
        if (done) {                                                        // :141-145
            if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = gbest;
            else { curve[clen] = gbest; clen += 1; }
        }
        S.w[b] = w; S.cur_index[b] = jn; S.fes[b] = fes; S.log_index[b] = lg; S.curve_len[b] = clen; S.done[b] = done ? 1 : 0;
    }
}

// =====================================================================================================
template <int D, int KIND>
static int rp_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxRlPsoState* s, const MbxRlPsoTape* tape, float* state, int B, cudaStream_t st) {
    auto kern = rlpso_init_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PopSmem<D, KIND>::kBytes));
        attr_done.set();
    }
    MbxRlPsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, PopSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, state);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int rp_launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxRlPsoState* s, const MbxRlPsoTape* tape, const float* action, float* state, float* reward,
                            int B, cudaStream_t st) {
    auto kern = rlpso_step_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RpSmem<D, KIND>::kBytes));
        attr_done.set();
    }
    MbxRlPsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, RpSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, state, reward);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_rl_pso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlPsoState* s, const MbxRlPsoTape* tape, float* state, int B,
                           cudaStream_t st) {
    MBX_DISPATCH(rp_launch_init_t, inst, consts, lane_inst, cfg, s, tape, state, B, st)
    mbx_set_error("mbx_rl_pso_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_rl_pso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlPsoState* s, const MbxRlPsoTape* tape, const float* action,
                           float* state, float* reward, int B, cudaStream_t st) {
    MBX_DISPATCH(rp_launch_step_t, inst, consts, lane_inst, cfg, s, tape, action, state, reward, B, st)
    mbx_set_error("mbx_rl_pso_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
