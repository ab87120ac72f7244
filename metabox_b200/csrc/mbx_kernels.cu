// mbx_kernels.cu -- stand-alone evaluation, Random_search and the fused RLEPSO (PSO family) kernels.
// One CTA of MBX_NT threads advances ONE lane (= one (instance, run) population).
#include <type_traits>
#include "mbx_eval.cuh"
#include "mbx_launch.h"

// =====================================================================================================
// K1/K3  problem.eval            grid (ceil(NP/RPC), B)
// =====================================================================================================
template <int D, int KIND>
struct EvalKernelCfg {
    static constexpr int kRowsPerCta = (KIND == MBX_KIND_PROTEIN) ? 32 : MBX_NT;
};

// Linear slope (F5), Rosenbrock (F8) and Schwefel (F20) -- and the noisy Rosenbrock ids built on F8 -- have no rotation: a candidate reads
// each of its 4 D vector constants once.  Staging the 15 KB constant block per CTA of 100 candidates (what eval_kernel does for the
// rotated functions) would move more bytes than the 4 D + 4 B per evaluation of the candidates themselves, and its shared-memory
// footprint allows 3 CTAs per SM.  For these lanes evaluation IS a streaming problem: eval_stream_kernel gives every candidate one
// thread, no shared memory, all 4 D bytes of the row requested before the first use (D / 2 independent 8-byte loads; consecutive threads
// = consecutive rows, every sector is consumed completely out of L1), constants through the read-only path (same address in the threads
// of a lane: broadcast).  The arithmetic is the per-candidate code of bbob_eval_cta (cases 5, 8, 20), operation for operation.
__host__ __device__ __forceinline__ bool eval_is_streaming(const int base) { return base == 5 || base == 8 || base == 20; }

#define MBX_STREAM_NT 256
template <int D, class TX>
__global__ void __launch_bounds__(MBX_STREAM_NT)
eval_stream_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   const TX* __restrict__ x, TX* __restrict__ out, int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen,
                   const double* __restrict__ noise_tape) {
    static_assert(D % 2 == 0, "rows are read as 8-byte (float) / 16-byte (double) pairs");
    // the CTA's MBX_STREAM_NT consecutive rows are one contiguous block of the input: fetched with 16-byte coalesced loads into shared memory
    // (a thread reading its own 4 D-byte row straight from global memory touches 32 different sectors per warp instruction: 4x the L1
    // transactions), then every thread takes its row from there
    extern __shared__ __align__(16) unsigned char stream_smem[];
    TX* tile = reinterpret_cast<TX*>(stream_smem);
    const size_t total = (size_t)B * NP, row0 = (size_t)blockIdx.x * MBX_STREAM_NT;
    const int nrows = (int)min((size_t)MBX_STREAM_NT, total - row0);
    {
        const TX* src = x + row0 * D;
        const int n = nrows * D;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            constexpr int PER = 16 / sizeof(TX);
            const int4* s4 = reinterpret_cast<const int4*>(src);
            int4* d4 = reinterpret_cast<int4*>(tile);
            for (int k = threadIdx.x; k < n / PER; k += MBX_STREAM_NT) d4[k] = __ldcs(s4 + k);
            for (int k = (n / PER) * PER + threadIdx.x; k < n; k += MBX_STREAM_NT) tile[k] = src[k];
        } else {
            for (int k = threadIdx.x; k < n; k += MBX_STREAM_NT) tile[k] = src[k];
        }
    }
    __syncthreads();
    if ((int)threadIdx.x >= nrows) return;
    const size_t gid = row0 + threadIdx.x;
    const int b = (int)(gid / NP), i = (int)(gid - (size_t)b * NP);
    const MbxInstance* gi = inst + lane_inst[b];
    const int base = gi->base;
    if (!eval_is_streaming(base)) return;
    // the row stays in its storage type (float: D registers, not 2 D) and is widened at each use
    using V2 = typename std::conditional<sizeof(TX) == 4, float2, double2>::type;
    V2 v[D / 2];
    {
        const V2* row = reinterpret_cast<const V2*>(tile + (size_t)threadIdx.x * D);
#pragma unroll
        for (int k = 0; k < D / 2; ++k) v[k] = row[k];
    }
#define xr_(d) ((double)(((d) & 1) ? v[(d) >> 1].y : v[(d) >> 1].x))
    const double* C = consts + gi->off;
    const double *cs = C, *cv0 = C + D + 2 * D * D, *cv1 = cv0 + D, *cv2 = cv1 + D;          // ConstView<D>
    const double ub = __ldg(&gi->ub), bias = __ldg(&gi->bias), bhc = __ldg(&gi->bh_coef);
    double f;
    if (base == 5) {                                                 // Linear slope (bbob_eval_cta case 5)
        f = 0.;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double xv = xr_(d);
            double z = xv;
            if (xv * __ldg(cs + d) > ub * ub) z = (z > 0. ? 1. : (z < 0. ? -1. : 0.)) * ub;
            const double sv = __ldg(cv0 + d);
            f += ub * fabs(sv) - z * sv;
        }
        f = f + bias;
    } else {
        double bh = 0.;
        if (bhc != 0. && base == 8) {                                // boundary handling (pen_row): base 8 only, like case 8's `+ bh`
            double p = 0.;
#pragma unroll
            for (int d = 0; d < D; ++d) { const double t = fabs(xr_(d)) - ub; if (t > 0.) p += t * t; }
            bh = (bhc == 1.) ? p : bhc * p;
        }
        if (base == 8) {                                             // Rosenbrock (case 8)
            const double sc0 = __ldg(&gi->sc[0]);
            f = 0.;
            double a = sc0 * (xr_(0) - __ldg(cs)) + 1.;
#pragma unroll
            for (int k = 0; k + 1 < D; ++k) {
                const double bq = sc0 * (xr_(k + 1) - __ldg(cs + k + 1)) + 1.;
                const double u = a * a - bq, w = a - 1.;
                f += 100. * (u * u) + w * w;
                a = bq;
            }
            f = f + bias + bh;
        } else {                                                     // Schwefel (case 20)
            double fs = 0., pen = 0., prev = 0.;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double z0 = __ldg(cv2 + d) * xr_(d);
                double z = z0;
                if (d > 0) z += 0.25 * (prev - __ldg(cv1 + d - 1));
                prev = z0;
                z = 100. * (__ldg(cv0 + d) * (z - __ldg(cv1 + d)) + __ldg(cv1 + d));
                fs += z * sin(sqrt(fabs(z)));
                const double t = fabs(z / 100.) - ub;
                if (t > 0.) pen += t * t;
            }
            f = 4.189828872724339 - 0.01 * (fs / (double)D) + 100. * pen + bias;
        }
    }
    const int noise = gi->noise;
    if (noise != MBX_NOISE_NONE) {                                   // same epilogue as cta_eval_tile
        MbxInstance h = *gi;
        double d0, d1, d2;
        if (noise_tape) {
            const double* nt = noise_tape + (size_t)b * 3 * NP;
            d0 = nt[i]; d1 = nt[NP + i]; d2 = nt[2 * NP + i];
        } else {
            Philox ph(seed);
            noise_draws(ph, noise, (uint32_t)i, gen, (uint32_t)b, MBX_TAG_NOISE, d0, d1, d2);
        }
        f = bbob_noise(h, f, d0, d1, d2);
    }
    if (subtract_opt && gi->has_optimum) f = f - __ldg(&gi->optimum);
    out[gid] = (TX)f;
#undef xr_
}

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
    if (KIND == MBX_KIND_BBOB && eval_is_streaming(gi->base)) return;          // those lanes belong to eval_stream_kernel
    stage_instance<D, KIND>(gi, consts, sh, sC);
    const double* cC = sC;
    const TX* src = x + ((size_t)b * NP + row0) * D;
    for (int k = threadIdx.x; k < n * D; k += MBX_NT) tile[k] = src[k];
    __syncthreads();
    Philox ph(seed);
    cta_eval_tile<D, KIND, TX, TX>(*sh, cC, consts + gi->off, tile, n, row0, scratch, outbuf, subtract_opt != 0, ph, gen,
                                   (uint32_t)b, MBX_TAG_NOISE, noise_tape ? noise_tape + (size_t)b * 3 * NP : nullptr, NP);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += MBX_NT) out[(size_t)b * NP + row0 + k] = outbuf[k];
}

template <int D, int KIND, class TX>
static int launch_eval_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const void* x, void* out,
                         int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen, const double* noise_tape,
                         cudaStream_t st, int paths) {
    constexpr int RPC = EvalKernelCfg<D, KIND>::kRowsPerCta;
    size_t smem = sizeof(MbxInstance) + sizeof(double) * (EvalTraits<D, KIND>::kConstDoubles + EvalTraits<D, KIND>::kScratchDoubles) +
                  sizeof(TX) * RPC * D;
    auto kern = eval_kernel<D, KIND, TX>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done.set();
    }
    if (KIND == MBX_KIND_BBOB && (paths & 1)) {                      // lanes of the no-rotation functions (the other kernel skips them)
        const size_t total = (size_t)B * NP;
        auto ks = eval_stream_kernel<D, TX>;
        const size_t sbytes = sizeof(TX) * MBX_STREAM_NT * D;
        static MbxPerDevice sattr_done;
        if (sbytes > 48 * 1024 && !sattr_done.get()) {
            MBX_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes));
            sattr_done.set();
        }
        ks<<<(unsigned)((total + MBX_STREAM_NT - 1) / MBX_STREAM_NT), MBX_STREAM_NT, sbytes, st>>>(
            inst, consts, lane_inst, (const TX*)x, (TX*)out, B, NP, subtract_opt, seed, gen, noise_tape);
        MBX_CUDA(cudaGetLastError());
    }
    if (KIND != MBX_KIND_BBOB || (paths & 2)) {
        dim3 grid((NP + RPC - 1) / RPC, B);
        kern<<<grid, MBX_NT, smem, st>>>(inst, consts, lane_inst, (const TX*)x, (TX*)out, NP, subtract_opt, seed, gen, noise_tape);
        MBX_CUDA(cudaGetLastError());
    }
    return MBX_OK;
}

// `kind` may carry the caller's promise about the lanes of the launch (include/metabox_b200.h: MBX_EVAL_ALL_STREAMING /
// MBX_EVAL_NONE_STREAMING): the kernel that would only skip every lane is then not launched at all (`paths` bit 0: eval_stream_kernel,
// bit 1: eval_kernel).
int mbx_launch_eval(int kind, int D, int x_is_f64, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                    const void* x, void* out, int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen,
                    const double* noise_tape, cudaStream_t st) {
    const int paths = (kind & MBX_EVAL_ALL_STREAMING) ? 1 : ((kind & MBX_EVAL_NONE_STREAMING) ? 2 : 3);
    kind &= 0xff;
#define MBX_EVAL_CASE(DD, KK)                                                                                          \
    if (D == DD && kind == KK)                                                                                         \
        return x_is_f64 ? launch_eval_t<DD, KK, double>(inst, consts, lane_inst, x, out, B, NP, subtract_opt, seed, gen, noise_tape, st, paths) \
                        : launch_eval_t<DD, KK, float>(inst, consts, lane_inst, x, out, B, NP, subtract_opt, seed, gen, noise_tape, st, paths);
#define MBX_EVAL_BBOB_(d, ...) MBX_EVAL_CASE(d, MBX_KIND_BBOB)
    MBX_FOR_BBOB_DIMS(MBX_EVAL_BBOB_, 0)
#undef MBX_EVAL_BBOB_
    MBX_EVAL_CASE(12, MBX_KIND_PROTEIN)
#undef MBX_EVAL_CASE
    mbx_set_error("mbx_eval: (kind=%d, dim=%d) is not compiled; BBOB dims {10, 20, 30} (+ -DMBX_DIM_EXTRA at build time), protein dim 12", kind, D);
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
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, launch_rs_t, inst, consts, lane_inst, cfg, tape_u, n_tape_iter, curve, curve_len, fes, B, st)
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return launch_rs_t<12, MBX_KIND_PROTEIN>(inst, consts, lane_inst, cfg, tape_u, n_tape_iter, curve, curve_len, fes, B, st);
    mbx_set_error("mbx_random_search: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
