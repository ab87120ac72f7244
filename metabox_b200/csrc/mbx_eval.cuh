// mbx_eval.cuh -- CTA-cooperative evaluation of a tile of candidates + constant staging.
// Shared by the stand-alone eval kernel and the fused optimizer kernels.
#pragma once
#include "mbx_bbob_cta.cuh"

template <int D, int KIND>
struct EvalTraits {
    static constexpr int kConstDoubles = (KIND == MBX_KIND_PROTEIN) ? MBX_PROT_DOUBLES : ConstView<D>::kDoubles;
    // scratch doubles a CTA of MBX_NT threads needs while evaluating
    static constexpr int kScratchDoubles =
        (KIND == MBX_KIND_PROTEIN) ? (MBX_NT / 32) * MBX_PROT_WARP_SCRATCH : (D + 1) * MBX_NT;
    // scratch doubles needed to evaluate n candidates (BBOB: the [n][D+1] tile of bbob_eval_cta)
    static __host__ __device__ constexpr int scratch_for(int n) {
        return (KIND == MBX_KIND_PROTEIN) ? (MBX_NT / 32) * MBX_PROT_WARP_SCRATCH : (D + 1) * n;
    }
};

// cooperative copy of the instance header + constants into shared memory
template <int D, int KIND>
__device__ __forceinline__ void stage_instance(const MbxInstance* __restrict__ inst, const double* __restrict__ consts,
                                               MbxInstance* sh, double* sC) {
    constexpr int n = EvalTraits<D, KIND>::kConstDoubles;
    const double* src = consts + inst->off;
    // every load of the thread is issued before the first store (n / 128 <= 32 values in registers): one exposed L2 / DRAM latency for the
    // whole block of constants instead of one per unrolled group of four
    constexpr int PER = (n + MBX_NT - 1) / MBX_NT;
    if (MBX_BDIM == MBX_NT && PER <= 32) {
        double r[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int k = MBX_TID + q * MBX_NT; r[q] = k < n ? __ldg(src + k) : 0.; }
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int k = MBX_TID + q * MBX_NT; if (k < n) sC[k] = r[q]; }
    } else {
        for (int k = MBX_TID; k < n; k += MBX_BDIM) sC[k] = __ldg(src + k);
    }
    const int* hs = reinterpret_cast<const int*>(inst);
    int* hd = reinterpret_cast<int*>(sh);
    for (int k = MBX_TID; k < (int)(sizeof(MbxInstance) / 4); k += MBX_BDIM) hd[k] = hs[k];
}

// Evaluate candidates [0,n) of a shared-memory tile (row-major [n][D], element type TX) and write
// cost = f(x) (- optimum) with noise into out (float32 if OUT32 else float64).  All threads of the
// CTA must call this (the BBOB path is CTA-cooperative and contains barriers; `scratch` needs
// EvalTraits::scratch_for(n) doubles; `zst` is unused).
//   noise_tape : per-candidate draws [3][stride] (d0,d1,d2) or nullptr -> Philox(seed, lane, gen, tag)
template <int D, int KIND, class TX, class TO>
__device__ __forceinline__ void cta_eval_tile(const MbxInstance& h, const double* sC, const double* gC,
                                              const TX* tile, int n, int row0, double* scratch, TO* out,
                                              bool subtract_opt, const Philox& ph, uint32_t gen, uint32_t lane,
                                              uint32_t tag, const double* noise_tape, int tape_stride,
                                              const unsigned char* only_mask = nullptr, const int zst = MBX_NT,
                                              void* gtile = nullptr, const int gtile_bytes = 0) {
    if (KIND == MBX_KIND_PROTEIN) {
        const int w = MBX_TID >> 5, l = MBX_TID & 31;
        double* ws = scratch + w * MBX_PROT_WARP_SCRATCH;
        for (int i = w; i < n; i += (MBX_NT >> 5)) {
            if (only_mask && !only_mask[i]) continue;
            double e;
            const unsigned short* plist = reinterpret_cast<const unsigned short*>(gC + MBX_PROT_DOUBLES);   // candidate pair list
            if (sizeof(TX) == 4) { XF32 X{reinterpret_cast<const float*>(tile) + i * D}; e = protein_energy_warp(sC, X, ws, l, plist, h.n_peaks, h.ub); }
            else { XF64 X{reinterpret_cast<const double*>(tile) + i * D}; e = protein_energy_warp(sC, X, ws, l, plist, h.n_peaks, h.ub); }
            if (l == 0) out[i] = (TO)e;
        }
    } else {
        // CTA-cooperative BBOB evaluation (mbx_bbob_cta.cuh): every thread participates, thread i receives f(x_i)
        const int i = MBX_TID;
        const bool act = i < n && !(only_mask && !only_mask[i]);
        double f = bbob_eval_cta<D, TX>(h, sC, gC, tile, n, scratch, only_mask, reinterpret_cast<float4*>(gtile), gtile_bytes);
        if (act) {
            if (h.noise != MBX_NOISE_NONE) {
                double d0, d1, d2;
                if (noise_tape) {
                    d0 = noise_tape[row0 + i]; d1 = noise_tape[tape_stride + row0 + i]; d2 = noise_tape[2 * tape_stride + row0 + i];
                } else {
                    noise_draws(ph, h.noise, (uint32_t)(row0 + i), gen, lane, tag, d0, d1, d2);
                }
                f = bbob_noise(h, f, d0, d1, d2);
            }
            if (subtract_opt && h.has_optimum) f = f - h.optimum;
            out[i] = (TO)f;
        }
    }
}

// fitness of ONE candidate (float32 tile `x`), noise applied.  ALL threads of the CTA call: the BBOB path is the
// CTA-cooperative evaluator (the D dimensions of the single trial are spread over D threads, mbx_bbob_cta.cuh), the
// protein path uses warp 0.  The result is valid in thread 0.
template <int D, int KIND>
__device__ __forceinline__ double cta_eval_one(const MbxInstance& h, const double* sC, const double* gC, const float* x,
                                              double* scratch, const Philox& ph, uint32_t idx, uint32_t gen, uint32_t lane,
                                              const double* noise_tape, int tape_stride) {
    double f = 0.;
    if (KIND == MBX_KIND_PROTEIN) {
        if (MBX_TID < 32) {
            XF32 X{x};
            f = protein_energy_warp(sC, X, scratch, MBX_TID & 31, reinterpret_cast<const unsigned short*>(gC + MBX_PROT_DOUBLES), h.n_peaks, h.ub);
        }
    } else {
        f = bbob_eval_cta<D, float>(h, sC, gC, x, 1, scratch, nullptr, nullptr, 0);
        if (MBX_TID == 0) {
            if (h.noise != MBX_NOISE_NONE) {
                double d0, d1, d2;
                if (noise_tape) { d0 = noise_tape[0]; d1 = noise_tape[tape_stride]; d2 = noise_tape[2 * tape_stride]; }
                else noise_draws(ph, h.noise, idx, gen, lane, MBX_TAG_NOISE, d0, d1, d2);
                f = bbob_noise(h, f, d0, d1, d2);
            }
            if (h.has_optimum) f = f - h.optimum;
        }
    }
    return f;
}

// =====================================================================================================
// shared-memory plan of the population kernels
// =====================================================================================================
template <int D, int KIND>
struct PopSmem {
    static constexpr int NPMAX = MBX_MAX_NP;
    static constexpr int kConst = EvalTraits<D, KIND>::kConstDoubles;
    static constexpr int kRegionA = (EvalTraits<D, KIND>::kScratchDoubles > NPMAX * D) ? EvalTraits<D, KIND>::kScratchDoubles : NPMAX * D;
    // byte offsets
    static constexpr size_t oHdr = 0;
    static constexpr size_t oConst = sizeof(MbxInstance);
    static constexpr size_t oA = oConst + sizeof(double) * kConst;           // px64 / eval scratch
    static constexpr size_t oPer = oA + sizeof(double) * kRegionA;           // 10 per-particle double arrays
    static constexpr size_t oGx = oPer + sizeof(double) * 10 * NPMAX;        // gbest_x as double [D]
    static constexpr size_t oXs = oGx + sizeof(double) * MBX_MAX_DIM;        // new positions f32 [NP*D]
    static constexpr size_t oF = oXs + sizeof(float) * NPMAX * D;            // 4 per-particle float arrays
    static constexpr size_t oI = oF + sizeof(float) * 4 * NPMAX;             // 2 per-particle int arrays
    static constexpr size_t oTgt = oI + sizeof(int) * 2 * NPMAX;             // FDR targets u8 [NP*D] + masks u8 [2*NP]
    static constexpr size_t oRed = oTgt + NPMAX * D + 2 * NPMAX;             // reduction scratch
    static constexpr size_t kBytes = ((oRed + 64 + 15) / 16) * 16;
};

