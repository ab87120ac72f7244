// mbx_kernels.cu -- stand-alone evaluation, Random_search and the fused RLEPSO (PSO family) kernels.
// One CTA of MBX_NT threads advances ONE lane (= one (instance, run) population).
#include "mbx_eval.cuh"
#include "mbx_launch.h"

// =====================================================================================================
// K1/K3  problem.eval            grid (ceil(NP/RPC), B)
// =====================================================================================================
template <int D, int KIND>
struct EvalKernelCfg {
    static constexpr int kRowsPerCta = (KIND == MBX_KIND_PROTEIN) ? 32 : MBX_NT;
};

template <int D, int KIND, class TX>
__global__ void __launch_bounds__(MBX_NT)
eval_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
            const TX* __restrict__ x, TX* __restrict__ out, int NP, int subtract_opt, uint64_t seed, uint32_t gen,
            const double* __restrict__ noise_tape) {
    constexpr int RPC = EvalKernelCfg<D, KIND>::kRowsPerCta;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw);
    double* sC = reinterpret_cast<double*>(smem_raw + sizeof(MbxInstance));
    double* scratch = sC + EvalTraits<D, KIND>::kConstDoubles;
    TX* tile = reinterpret_cast<TX*>(scratch + EvalTraits<D, KIND>::kScratchDoubles);
    __shared__ TX outbuf[RPC];

    const int b = blockIdx.y;
    const int row0 = blockIdx.x * RPC;
    const int n = min(RPC, NP - row0);
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, sh, sC);
    const TX* src = x + ((size_t)b * NP + row0) * D;
    for (int k = threadIdx.x; k < n * D; k += MBX_NT) tile[k] = src[k];
    __syncthreads();
    Philox ph(seed);
    cta_eval_tile<D, KIND, TX, TX>(*sh, sC, consts + gi->off, tile, n, row0, scratch, outbuf, subtract_opt != 0, ph, gen,
                                   (uint32_t)b, MBX_TAG_NOISE, noise_tape ? noise_tape + (size_t)b * 3 * NP : nullptr, NP);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += MBX_NT) out[(size_t)b * NP + row0 + k] = outbuf[k];
}

template <int D, int KIND, class TX>
static int launch_eval_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const void* x, void* out,
                         int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen, const double* noise_tape,
                         cudaStream_t st) {
    constexpr int RPC = EvalKernelCfg<D, KIND>::kRowsPerCta;
    size_t smem = sizeof(MbxInstance) + sizeof(double) * (EvalTraits<D, KIND>::kConstDoubles + EvalTraits<D, KIND>::kScratchDoubles) +
                  sizeof(TX) * RPC * D;
    auto kern = eval_kernel<D, KIND, TX>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done.set();
    }
    dim3 grid((NP + RPC - 1) / RPC, B);
    kern<<<grid, MBX_NT, smem, st>>>(inst, consts, lane_inst, (const TX*)x, (TX*)out, NP, subtract_opt, seed, gen, noise_tape);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

int mbx_launch_eval(int kind, int D, int x_is_f64, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                    const void* x, void* out, int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen,
                    const double* noise_tape, cudaStream_t st) {
#define MBX_EVAL_CASE(DD, KK)                                                                                          \
    if (D == DD && kind == KK)                                                                                         \
        return x_is_f64 ? launch_eval_t<DD, KK, double>(inst, consts, lane_inst, x, out, B, NP, subtract_opt, seed, gen, noise_tape, st) \
                        : launch_eval_t<DD, KK, float>(inst, consts, lane_inst, x, out, B, NP, subtract_opt, seed, gen, noise_tape, st);
    MBX_EVAL_CASE(10, MBX_KIND_BBOB)
    MBX_EVAL_CASE(30, MBX_KIND_BBOB)
    MBX_EVAL_CASE(12, MBX_KIND_PROTEIN)
#undef MBX_EVAL_CASE
    mbx_set_error("mbx_eval: (kind=%d, dim=%d) is not compiled; BBOB dims {10,30}, protein dim 12", kind, D);
    return MBX_EUNSUP;
}

// =====================================================================================================
// K7  Random_search.run_episode (optimizer/random_search.py:17-59): whole episode per lane
// =====================================================================================================
template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT)
random_search_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts,
                     const int32_t* __restrict__ lane_inst, MbxRunConfig cfg, const double* __restrict__ tape_u, int n_tape_iter,
                     float* __restrict__ curve, int32_t* __restrict__ curve_len, int32_t* __restrict__ fes_out) {
    using L = PopSmem<D, KIND>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(smem_raw + L::oHdr);
    double* sC = reinterpret_cast<double*>(smem_raw + L::oConst);
    double* scratch = reinterpret_cast<double*>(smem_raw + L::oA);
    float* xs = reinterpret_cast<float*>(smem_raw + L::oXs);
    float* cost = reinterpret_cast<float*>(smem_raw + L::oF);
    float* redv = reinterpret_cast<float*>(smem_raw + L::oRed);
    int* redi = reinterpret_cast<int*>(smem_raw + L::oRed + 32);

    const int b = blockIdx.x, NP = cfg.NP;
    const MbxInstance* gi = inst + lane_inst[b];
    stage_instance<D, KIND>(gi, consts, sh, sC);
    __syncthreads();
    const MbxInstance& h = *sh;
    Philox ph(cfg.seed);
    const double lb = h.lb, ub = h.ub;
    int fes = 0, log_index = 1, clen = 0;
    float gbest = 0.f;
    float* mycurve = curve + (size_t)b * MBX_CURVE_LEN;
    for (int it = 0;; ++it) {
        for (int e = threadIdx.x; e < NP * D; e += MBX_NT) {
            double u;
            if (cfg.use_tape) u = tape_u[(((size_t)b * n_tape_iter + it) * NP) * D + e];
            else u = u01(ph((uint32_t)e, (uint32_t)it, (uint32_t)(cfg.lane_base + b), MBX_TAG_INIT).x);
            xs[e] = (float)(lb + (ub - lb) * u);
        }
        __syncthreads();
        cta_eval_tile<D, KIND, float, float>(h, sC, consts + gi->off, xs, NP, 0, scratch, cost, true, ph, (uint32_t)it,
                                             (uint32_t)(cfg.lane_base + b), MBX_TAG_NOISE, nullptr, 0);
        __syncthreads();
        float v = threadIdx.x < NP ? cost[threadIdx.x] : INFINITY;
        int ix = threadIdx.x;
        block_argmin<MBX_NT>(v, ix, redv, redi);
        fes += NP;
        bool done;
        if (it == 0) {                                   // __reset: random_search.py:17-22
            gbest = v;
            if (threadIdx.x == 0) mycurve[0] = gbest;
            clen = 1;
            done = false;
            // the reference's first loop iteration happens after reset; nothing else here
        } else {
            if (gbest > v) gbest = v;
            if (fes >= log_index * cfg.log_interval) {   // random_search.py:43-45
                log_index += 1;
                if (threadIdx.x == 0 && clen < MBX_CURVE_LEN) mycurve[clen] = gbest;
                clen = min(clen + 1, MBX_CURVE_LEN);
            }
            done = h.has_optimum ? ((double)gbest <= 1e-8 || fes >= cfg.maxFEs) : (fes >= cfg.maxFEs);
            if (done) {                                  // random_search.py:52-57
                if (clen >= cfg.n_logpoint + 1) { if (threadIdx.x == 0) mycurve[clen - 1] = gbest; }
                else { if (threadIdx.x == 0) mycurve[clen] = gbest; clen += 1; }
            }
        }
        __syncthreads();
        if (done) break;
    }
    if (threadIdx.x == 0) { curve_len[b] = clen; fes_out[b] = fes; }
}

template <int D, int KIND>
static int launch_rs_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                       const double* tape_u, int n_tape_iter, float* curve, int32_t* curve_len, int32_t* fes, int B,
                       cudaStream_t st) {
    auto kern = random_search_kernel<D, KIND>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PopSmem<D, KIND>::kBytes));
        attr_done.set();
    }
    kern<<<B, MBX_NT, PopSmem<D, KIND>::kBytes, st>>>(inst, consts, lane_inst, *cfg, tape_u, n_tape_iter, curve, curve_len, fes);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

int mbx_launch_rs(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                  const double* tape_u, int n_tape_iter, float* curve, int32_t* curve_len, int32_t* fes, int B, cudaStream_t st) {
    if (cfg->dim == 10 && kind == MBX_KIND_BBOB) return launch_rs_t<10, MBX_KIND_BBOB>(inst, consts, lane_inst, cfg, tape_u, n_tape_iter, curve, curve_len, fes, B, st);
    if (cfg->dim == 30 && kind == MBX_KIND_BBOB) return launch_rs_t<30, MBX_KIND_BBOB>(inst, consts, lane_inst, cfg, tape_u, n_tape_iter, curve, curve_len, fes, B, st);
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return launch_rs_t<12, MBX_KIND_PROTEIN>(inst, consts, lane_inst, cfg, tape_u, n_tape_iter, curve, curve_len, fes, B, st);
    mbx_set_error("mbx_random_search: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
