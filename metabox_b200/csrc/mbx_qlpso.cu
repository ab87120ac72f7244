// mbx_qlpso.cu -- fused QLPSO step (PSO family, ONE particle per step; the action picks the size of the ring neighbourhood):
// neighbourhood best over the CURRENT costs -> velocity of the current particle from two scalar draws -> position, clipped
// -> swarm diversity (mean distance to the centroid, in numpy's reduction order so that the float64 value and with it the
// integer reward are bit-identical) -> fitness of the single candidate (CTA-cooperative) -> reward in {2, 1, 0, -2} ->
// pbest row / gbest cost / per-particle Q-state -> cost-curve log -> Q-state of the NEXT particle.
// One CTA per lane; state float32 in HBM / float64 in registers; the diversity scalar stays float64.
// Reference: optimizer/qlpso_optimizer.py:7-16 (cal_reward), :45-46 (diversity), :48-72 (velocity), :74-89
// (init_population), :91-126 (update); fixed configuration NP = 30, C = 1.49618, W = 0.729844 (:22-25).
// numpy twin: oracle/qlpso.py.
#include "mbx_eval.cuh"
#include "mbx_launch.h"

#define QL_C 1.49618
#define QL_W 0.729844

// numpy's pairwise summation for n <= 128 contiguous float64 (8 accumulators, then the tail one by one); GET(k) -> element k
template <typename F>
__device__ __forceinline__ double np_sum128(const int n, F get) {
    if (n < 8) {
        double r = 0.;
        for (int k = 0; k < n; ++k) r += get(k);
        return r;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = get(k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] += get(i + k);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += get(i);
    return res;
}

// np.mean(np.sqrt(np.sum(np.square(pop - np.mean(pop, 0)), 1)))  (:45-46) over the float32 positions xs [NP*D] in shared memory.
// cen [D], dist [NP] shared scratch.  Contains CTA barriers; the result is valid in thread 0.
template <int D>
__device__ __forceinline__ double ql_diversity(const float* xs, const int NP, double* cen, double* dist) {
    const int tid = threadIdx.x;
    if (tid < D) {                                                 // axis-0 reduction: row after row
        double c = 0.;
        for (int i = 0; i < NP; ++i) c += (double)xs[i * D + tid];
        cen[tid] = c / (double)NP;
    }
    __syncthreads();
    if (tid < NP) {
        const float* xr = xs + tid * D;
        const double s = np_sum128(D, [&](int k) { const double dv = (double)xr[k] - cen[k]; return dv * dv; });
        dist[tid] = sqrt(s);
    }
    __syncthreads();
    double r = 0.;
    if (tid == 0) r = np_sum128(NP, [&](int k) { return dist[k]; }) / (double)NP;
    return r;
}

// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
qlpso_init_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxQlpsoState S, MbxQlpsoTape T, int32_t* __restrict__ state) {
    using L = PopSmem<D, KIND>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw + L::oHdr);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oA);
    float* xs = reinterpret_cast<float*>(smem_raw + L::oXs);
    float* sc = reinterpret_cast<float*>(smem_raw + L::oF);
    float* redv = reinterpret_cast<float*>(smem_raw + L::oRed);
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
        const double up = cfg.use_tape ? T.init_pos_u[base + e] : u01(ph((uint32_t)e, MBX_GEN_INIT, lane, MBX_TAG_INIT).x);
        const float p = (float)(up * (ub - lb) + lb);                      // :75
        xs[e] = p;
        S.x[base + e] = p; S.pbest_x[base + e] = p; S.v[base + e] = 0.f;
    }
    __syncthreads();
    const double div = ql_diversity<D>(xs, NP, scratch, scratch + D);       // :78 (before the evaluation scratch is used)
    __syncthreads();
    cta_eval_tile<D, KIND, float, float>(h, sC, consts + gi->off, xs, NP, 0, scratch, sc, true, ph, MBX_GEN_INIT, lane, MBX_TAG_NOISE,
                                         cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 * NP : nullptr, NP);
    __syncthreads();
    float v = INFINITY;
    if (tid < NP) {
        v = sc[tid]; S.cost[pb + tid] = v;
        S.q_state[pb + tid] = cfg.use_tape ? T.init_state[pb + tid] : (int32_t)(ph((uint32_t)tid, MBX_GEN_INIT, lane, MBX_TAG_QLPSO).x & 3u);   // :88
    }
    const float best = -block_max<MBX_NT>(tid < NP ? -v : -INFINITY, redv);
    if (tid == 0) {
        S.gbest_cost[b] = best; S.diversity[b] = div; S.pointer[b] = 0;
        S.fes[b] = NP; S.log_index[b] = 1; S.curve[(size_t)b * MBX_CURVE_LEN] = best; S.curve_len[b] = 1; S.done[b] = 0;
        state[b] = cfg.use_tape ? T.init_state[pb] : (int32_t)(ph(0u, MBX_GEN_INIT, lane, MBX_TAG_QLPSO).x & 3u);
    }
}

// =====================================================================================================
struct QlLayout {
    size_t oConst, oScr, oXs, oCost, oCen, oDist, oTrial, oMisc, bytes;
    __host__ __device__ QlLayout(int NP, int D, int kConst, int kScratch) {
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oScr = oConst + sizeof(double) * kConst;
        oCen = oScr + sizeof(double) * kScratch;
        oDist = oCen + sizeof(double) * D;
        oXs = oDist + sizeof(double) * NP;
        oCost = oXs + sizeof(float) * NP * D;
        oTrial = oCost + sizeof(float) * ((NP + 1) & ~1);
        oMisc = (oTrial + sizeof(float) * D + 15) & ~(size_t)15;
        bytes = oMisc + 64;
    }
};
template <int D, int KIND>
__host__ __device__ inline QlLayout ql_layout(int NP) {
    return QlLayout(NP, D, (KIND == MBX_KIND_PROTEIN) ? EvalTraits<D, KIND>::kConstDoubles : 0,   // BBOB constants stay in global
                    (KIND == MBX_KIND_PROTEIN) ? MBX_PROT_WARP_SCRATCH : D + 1);
}

template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT, 8)
qlpso_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxQlpsoState S, MbxQlpsoTape T, const int32_t* __restrict__ action, int32_t* __restrict__ state,
                  float* __restrict__ reward) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }      // finished lane: state frozen, reward 0
    const int NP = cfg.NP, tid = threadIdx.x;
    const QlLayout L = ql_layout<D, KIND>(NP);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw);
    double* sC = reinterpret_cast<double*>(smem_raw + L.oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L.oScr);
    double* cen = reinterpret_cast<double*>(smem_raw + L.oCen);
    double* dist = reinterpret_cast<double*>(smem_raw + L.oDist);
    float* xs = reinterpret_cast<float*>(smem_raw + L.oXs);
    float* cs = reinterpret_cast<float*>(smem_raw + L.oCost);
    float* trial = reinterpret_cast<float*>(smem_raw + L.oTrial);
    double* md = reinterpret_cast<double*>(smem_raw + L.oMisc);            // [0] new diversity
    int* mi = reinterpret_cast<int*>(smem_raw + L.oMisc + 16);             // [0] the particle improved on its previous cost
    const MbxInstance* gi = inst + lane_inst[b];
    const double* cC = sC;
    if (KIND == MBX_KIND_PROTEIN) stage_instance<D, KIND>(gi, consts, sh, sC);
    else {
        const int* hs = reinterpret_cast<const int*>(gi);
        int* hd = reinterpret_cast<int*>(sh);
        for (int k = tid; k < (int)(sizeof(MbxInstance) / 4); k += MBX_NT) hd[k] = hs[k];
        cC = consts + gi->off;
    }
    const int i = S.pointer[b], fes0 = S.fes[b], act = action[b];
    const size_t pb = (size_t)b * NP, base = pb * D, row = base + (size_t)i * D;
    for (int e = tid; e < NP * D; e += MBX_NT) xs[e] = S.x[base + e];
    if (tid < NP) cs[tid] = S.cost[pb + tid];
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    __syncthreads();
    const MbxInstance& h = *sh;
    const double lb = h.lb, ub = h.ub;
    // ---- ring-neighbourhood best of the current costs (:53-70): first strict minimum, every thread on its own -------------------
    int nb = i;
    {
        const int k = act == 0 ? 4 : act == 1 ? 8 : act == 2 ? 16 : act == 3 ? 30 : 0;
        float bc = INFINITY;
        for (int j = -(k / 2); j <= k / 2; ++j) {
            const int m = (((i + j) % NP) + NP) % NP;
            const float c = cs[m];
            if (c < bc) { bc = c; nb = m; }
        }
    }
    // ---- velocity + position of particle i (:71-72, :92-95); no velocity clamp in this backbone ---------------------------------
    if (tid < D) {
        const int d = tid;
        double r1, r2;
        if (cfg.use_tape) { r1 = T.rand2[2 * b]; r2 = T.rand2[2 * b + 1]; }
        else { const uint4 w = ph(0u, (uint32_t)fes0, lane, MBX_TAG_QLPSO); r1 = u01(w.x); r2 = u01(w.y); }
        const double x = (double)xs[i * D + d], v = (double)S.v[row + d], pbx = (double)S.pbest_x[row + d];
        const double nbx = (double)xs[nb * D + d];
        const double nv = QL_W * v + QL_C * r1 * (nbx - x) + QL_C * r2 * (pbx - x);
        double nx = x + nv;
        nx = nx < lb ? lb : nx; nx = nx > ub ? ub : nx;
        const float nx32 = (float)nx;
        trial[d] = nx32;
        S.x[row + d] = nx32; S.v[row + d] = (float)nv;
    }
    __syncthreads();
    if (tid < D) xs[i * D + tid] = trial[tid];                              // (after every thread has read the neighbour row)
    __syncthreads();
    // ---- diversity of the moved swarm (:103) ---------------------------------------------------------------------------------------
    {
        const double dn = ql_diversity<D>(xs, NP, cen, dist);
        if (tid == 0) md[0] = dn;
    }
    // ---- fitness of the single candidate (:98-101) -----------------------------------------------------------------------------
    const double f = cta_eval_one<D, KIND>(h, cC, consts + gi->off, trial, scratch, ph, 0u, (uint32_t)fes0, lane,
                                           cfg.use_tape && T.noise ? T.noise + (size_t)b * 3 : nullptr, 1);
    if (tid == 0) {
        const float f_new = (float)f, f_old = cs[i];
        const double d_new = md[0], d_old = S.diversity[b];
        const bool better = f_new < f_old, wider = d_new > d_old;
        reward[b] = better ? (wider ? 2.f : 1.f) : (wider ? 0.f : -2.f);   // cal_reward (:7-16)
        mi[0] = better ? 1 : 0;
        S.cost[pb + i] = f_new; S.diversity[b] = d_new;
        float gbest = S.gbest_cost[b];
        if (f_new < gbest) { gbest = f_new; S.gbest_cost[b] = gbest; }     // min(gbest, cost.min()) (:108)
        S.q_state[pb + i] = act;                                           // :111
        const int nxt = (i + 1) % NP;
        const int fes = fes0 + 1;
        int lg = S.log_index[b], clen = S.curve_len[b];
        float* curve = S.curve + (size_t)b * MBX_CURVE_LEN;
        if (fes >= lg * cfg.log_interval) {                                // :114-116
            lg += 1;
            if (clen < MBX_CURVE_LEN) curve[clen] = gbest;
            clen = min(clen + 1, MBX_CURVE_LEN);
        }
        const bool done = h.has_optimum ? (fes >= cfg.maxFEs || (double)gbest <= 1e-8) : (fes >= cfg.maxFEs);
        if (done) {                                                        // :122-125
            if (clen >= cfg.n_logpoint + 1) curve[clen - 1] = gbest;
            else { curve[clen] = gbest; clen += 1; }
        }
        state[b] = (nxt == i) ? act : S.q_state[pb + nxt];
        S.pointer[b] = nxt; S.fes[b] = fes; S.log_index[b] = lg; S.curve_len[b] = clen; S.done[b] = done ? 1 : 0;
    }
    __syncthreads();
    if (mi[0] && tid < D) S.pbest_x[row + tid] = trial[tid];                // :109-110
}

// =====================================================================================================
template <int D, int KIND>
static int ql_launch_init_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxQlpsoState* s, const MbxQlpsoTape* tape, int32_t* state, int B, cudaStream_t st) {
    auto kern = qlpso_init_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PopSmem<D, KIND>::kBytes));
        attr_done.set();
    }
    MbxQlpsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, PopSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, state);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND>
static int ql_launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                            const MbxQlpsoState* s, const MbxQlpsoTape* tape, const int32_t* action, int32_t* state, float* reward,
                            int B, cudaStream_t st) {
    auto kern = qlpso_step_kernel<D, KIND>;
    const size_t bytes = ql_layout<D, KIND>(cfg->NP).bytes;
    static size_t attr_bytes = 0;
    if (bytes > attr_bytes) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        attr_bytes = bytes;
    }
    MbxQlpsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, action, state, reward);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return FN<12, MBX_KIND_PROTEIN>(__VA_ARGS__);

int mbx_launch_qlpso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxQlpsoState* s, const MbxQlpsoTape* tape, int32_t* state, int B,
                          cudaStream_t st) {
    MBX_DISPATCH(ql_launch_init_t, inst, consts, lane_inst, cfg, s, tape, state, B, st)
    mbx_set_error("mbx_qlpso_init: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_launch_qlpso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxQlpsoState* s, const MbxQlpsoTape* tape, const int32_t* action,
                          int32_t* state, float* reward, int B, cudaStream_t st) {
    MBX_DISPATCH(ql_launch_step_t, inst, consts, lane_inst, cfg, s, tape, action, state, reward, B, st)
    mbx_set_error("mbx_qlpso_step: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
