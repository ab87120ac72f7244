// mbx_deddqn.cu -- DE-DDQN: init kernel and the per-step kernel (one launch per function evaluation, action from outside) + launchers.
// The step's device code lives in mbx_deddqn.cuh (shared with the persistent kernel of mbx_deddqn_run.cu).
#include "mbx_deddqn.cuh"

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
    DdLane Ln{};
    Ln.X = S.X + base; Ln.Xg = S.X_gbest + (size_t)b * D; Ln.Xp = S.X_prebest + (size_t)b * D;
    Ln.n_tot = S.n_tot + (size_t)b * 4 * DD_GEN_MAX; Ln.n_succ = S.n_succ + (size_t)b * 16 * DD_GEN_MAX;
    Ln.om_sum = S.om_sum + (size_t)b * 16 * DD_GEN_MAX; Ln.om_max = S.om_max + (size_t)b * 16 * DD_GEN_MAX;
    Ln.sc = sc; Ln.swin = swin; Ln.red = red; Ln.feat = feature + (size_t)b * DD_NFEAT;        // (value-initialised: sorted = nullptr)
    dd_features<D>(cfg, Ln, 0, 0, 0, ix, 0, NP, v, worst, v, sr, lb, ub);
}

// =====================================================================================================
#ifndef MBX_DEDDQN_MIN_CTAS
#define MBX_DEDDQN_MIN_CTAS 8        // compile-time knob: minimum CTAs per SM of the step kernel (A/B builds: python -m metabox_b200.build --variant)
#endif
template <int D, int KIND>
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
    const size_t pb = (size_t)b * NP;
    DdLane Ln;
    Ln.X = S.X + (size_t)b * NP * D; Ln.Xg = S.X_gbest + (size_t)b * D; Ln.Xp = S.X_prebest + (size_t)b * D;
    Ln.n_tot = S.n_tot + (size_t)b * 4 * DD_GEN_MAX; Ln.n_succ = S.n_succ + (size_t)b * 16 * DD_GEN_MAX;
    Ln.om_sum = S.om_sum + (size_t)b * 16 * DD_GEN_MAX; Ln.om_max = S.om_max + (size_t)b * 16 * DD_GEN_MAX;
    Ln.curve = S.curve + (size_t)b * MBX_CURVE_LEN;
    Ln.cost_wt = S.cost + pb; Ln.win_wt = S.win + (size_t)b * DD_W * DD_WS; Ln.r_wt = S.r + (size_t)b * 5;
    Ln.sc = reinterpret_cast<float*>(smem_raw + L::oCost);
    Ln.sorted = nullptr;
    Ln.swin = reinterpret_cast<double*>(smem_raw + L::oWin);
    Ln.red = reinterpret_cast<double*>(smem_raw + L::oRedD);
    Ln.scratch = reinterpret_cast<double*>(smem_raw + L::oScr);
    Ln.trial = reinterpret_cast<float*>(smem_raw + L::oTrial);
    Ln.mf = reinterpret_cast<float*>(smem_raw + L::oMisc);
    Ln.mi = reinterpret_cast<int*>(smem_raw + L::oMisc + 32);
    Ln.feat = feature + (size_t)b * DD_NFEAT;
    DdRegs R;
    R.p = ints[DD_I_POINTER]; R.gen = ints[DD_I_GEN]; R.stag = ints[DD_I_STAG]; R.prow = ints[DD_I_PROW]; R.rebound = ints[DD_I_REBOUND];
    R.wlen = ints[DD_I_WLEN]; R.wseq = ints[DD_I_WSEQ];
    R.fes = S.fes[b]; R.log_index = S.log_index[b]; R.curve_len = S.curve_len[b];
    R.cgb = S.c3[(size_t)b * 3]; R.cgw = S.c3[(size_t)b * 3 + 1]; R.cpb = S.c3[(size_t)b * 3 + 2];
    R.done = false;
    if (tid < NP) Ln.sc[tid] = S.cost[pb + tid];
    if (tid < 5) Ln.mi[tid] = S.r[(size_t)b * 5 + tid];
    for (int k = tid; k < R.wlen * DD_WS; k += MBX_NT) Ln.swin[k] = S.win[(size_t)b * DD_W * DD_WS + k];
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    // (the barrier that publishes the staged header, costs, window and indices is the first one inside dd_step_core)
    const float rew = dd_step_core<D, KIND>(cfg, *sh, cC, consts + gi->off, Ln, R, a, ph, lane,
                                            cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 : nullptr,
                                            cfg.use_tape ? T.r5 + (size_t)b * 5 : nullptr, dbg ? dbg + (size_t)b * (D + 1) : nullptr);
    if (tid == 0) {
        reward[b] = rew;
        ints[DD_I_POINTER] = R.p; ints[DD_I_GEN] = R.gen; ints[DD_I_STAG] = R.stag; ints[DD_I_PROW] = R.prow;
        ints[DD_I_REBOUND] = R.rebound; ints[DD_I_WLEN] = R.wlen; ints[DD_I_WSEQ] = R.wseq;
        S.c3[(size_t)b * 3] = R.cgb; S.c3[(size_t)b * 3 + 1] = R.cgw;
        S.log_index[b] = R.log_index; S.curve_len[b] = R.curve_len; S.fes[b] = R.fes; S.done[b] = R.done ? 1 : 0;
    }
}

// =====================================================================================================
template <int D, int KIND>
static int dd_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxDdqnState* s, const MbxDdqnTape* tape, float* feature, int B, cudaStream_t st) {
    auto kern = deddqn_init_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PopSmem<D, KIND>::kBytes));
        attr_done.set();
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
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DdqnSmem<D, KIND>::kBytes));
        attr_done.set();
    }
    MbxDdqnTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, DdqnSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, feature, reward, dbg);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
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
