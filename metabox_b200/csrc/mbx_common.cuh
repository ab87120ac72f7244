// mbx_common.cuh -- shared device helpers: Philox4x32-10, block reductions, smem staging.
// Compiled with -fmad=false: every a*b+c written in source stays two IEEE operations (bit-identical
// to the numpy ufunc chain of the reference); fused multiply-adds are requested explicitly with fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/metabox_b200.h"

// Threads per lane (a lane = one (instance, run) population).  The population kernels give a lane a whole CTA of 128 threads; a translation
// unit may choose another group size BEFORE including the shared device code (the persistent DE-DDQN kernel, mbx_deddqn_run.cu: one WARP per
// lane, so that many lanes are resident per SM and their serial per-step chains overlap).
#ifndef MBX_NT
#define MBX_NT 128
#endif

// Lane groups.  A translation unit whose kernels advance several lanes per CTA defines MBX_LANE_GROUPS > 1 before including the shared
// device code: thread ids become group-local and the barriers inside the cooperative evaluator / reductions become barriers of the group
// (__syncwarp for one-warp groups, else the named barrier 1 + group), so that lanes of one CTA may take different paths (different
// functions, finished lanes).  With the defaults the macros are the plain built-ins.
#ifndef MBX_LANE_GROUPS
#define MBX_LANE_GROUPS 1
#endif
#if MBX_LANE_GROUPS == 1
#define MBX_TID ((int)threadIdx.x)
#define MBX_BDIM ((int)blockDim.x)
#define MBX_SYNC() __syncthreads()
#else
static_assert(MBX_NT % 32 == 0 && (MBX_NT & (MBX_NT - 1)) == 0, "lane groups are whole warps, a power of two of threads");
#define MBX_TID ((int)(threadIdx.x & (MBX_NT - 1)))
#define MBX_BDIM MBX_NT
#if MBX_NT == 32
#define MBX_SYNC() __syncwarp()
#else
#define MBX_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / MBX_NT)), "n"(MBX_NT) : "memory")
#endif
#endif

// ---- draw-group tags (numpy twin: oracle/rng.py) -------------------------------------------------
#define MBX_TAG_INIT 1u
#define MBX_TAG_PD 2u
#define MBX_TAG_P 3u
#define MBX_TAG_REINIT 4u
#define MBX_TAG_NOISE 5u
#define MBX_TAG_NOISE2 6u
#define MBX_TAG_LDE_P 7u
#define MBX_TAG_LDE_PD 8u
#define MBX_TAG_LDE_REJ 9u
#define MBX_TAG_DDQN 10u
#define MBX_TAG_DDQN_D 11u
#define MBX_TAG_GLEET_P 12u
#define MBX_TAG_RLPSO 13u
#define MBX_TAG_QLPSO 14u
#define MBX_GEN_INIT 0xFFFFFFFFu

// Phase profiling build (python -m metabox_b200.build --profile -> libmbx_prof.so, never the product library): thread 0 of
// every CTA accumulates clock64() deltas per phase into g_prof (one copy per translation unit; mbx_prof_read() in
// mbx_rlepso.cu reads that unit's copy).
#ifdef MBX_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_INIT long long tp_ = clock64();
#define PROF_MARK(k)                                                                   \
    do {                                                                               \
        if (threadIdx.x == 0) {                                                        \
            long long t_ = clock64();                                                  \
            atomicAdd(&g_prof[k], (unsigned long long)(t_ - tp_));                     \
            tp_ = t_;                                                                  \
        }                                                                              \
    } while (0)
#define PROF_ADD(k, v) do { if (threadIdx.x == 0) atomicAdd(&g_prof[k], (unsigned long long)(v)); } while (0)
#define PROF_ADD_ANY(k, v) atomicAdd(&g_prof[k], (unsigned long long)(v))
#else
#define PROF_INIT
#define PROF_MARK(k)
#define PROF_ADD(k, v)
#define PROF_ADD_ANY(k, v)
#endif

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    // counter = (index, generation, lane, tag)
    __device__ __forceinline__ uint4 operator()(uint32_t idx, uint32_t gen, uint32_t lane, uint32_t tag) const {
        uint32_t c0 = idx, c1 = gen, c2 = lane, c3 = tag, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// w / 2^32.  The integer -> float64 conversion instruction (I2F.F64.U32) runs at a quarter of the FP64 rate; 2^52 + w is exact in float64
// (w sits in the low mantissa bits), so one DADD gives the same value bit for bit.
__device__ __forceinline__ double u32_to_double(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)w) - 4503599627370496.0;
#else
    return (double)w;
#endif
}
__device__ __forceinline__ double u01(uint32_t w) { return u32_to_double(w) * (1.0 / 4294967296.0); }
__device__ __forceinline__ int bounded(uint32_t w, int high) { return (int)__umulhi(w, (uint32_t)high); }
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, double& n0, double& n1) {
    double ua = (u32_to_double(wa) + 1.0) * (1.0 / 4294967296.0);
    double ub = u32_to_double(wb) * (1.0 / 4294967296.0);
    double r = sqrt(-2.0 * log(ua));
    double t = 6.283185307179586 * ub;
    double s, c;
    sincos(t, &s, &c);
    n0 = r * c; n1 = r * s;
}

// Order-preserving 32-bit key of a float32 cost for an ascending sort: -0 == +0, NaN last (np.argsort puts NaN at the end).
__device__ __forceinline__ unsigned order_key(const float f) {
    if (f != f) return 0xFFFFFFFFu;
    const unsigned u = __float_as_uint(f + 0.f);                       // -0 + 0 = +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Stable ascending rank of one float32 cost per thread (threads >= n do not take part in the order) by a bitonic network over the 128
// threads of a CTA: the items are (order_key, thread index) in 64 bits, so the order is total and ties resolve by index like a stable
// sort.  25 of the 28 compare-exchange stages are warp shuffles; the 3 with a partner in another warp go through `xch` (128 x 8 bytes of
// shared memory that is dead at the call site).  ~10 instructions per stage instead of ~14 per PAIR of the all-pairs count it replaces.
// All 128 threads must call.  Returns the index of the item that ends at position threadIdx.x (valid for threadIdx.x < n).
__device__ __forceinline__ int block_sort_order_128(const float f, const int n, unsigned long long* xch) {
    const int t = threadIdx.x;
    unsigned long long v = t < n ? ((unsigned long long)order_key(f) << 32) | (unsigned)t : ~0ull;
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long o;
            if (j < 32) o = __shfl_xor_sync(0xffffffffu, v, j);
            else { xch[t] = v; __syncthreads(); o = xch[t ^ j]; __syncthreads(); }
            const bool takemin = ((t & j) == 0) == ((t & k) == 0);   // lower item of an ascending pair / upper item of a descending one
            v = ((o < v) == takemin) ? o : v;
        }
    }
    return (int)(unsigned)(v & 0xFFFFFFFFull);
}

// (value, index) lexicographic minimum == np.argmin first-occurrence rule
__device__ __forceinline__ void argmin_pair(float& v, int& i, float ov, int oi) {
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void warp_argmin(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmin_pair(v, i, ov, oi);
    }
}
// block-wide argmin over NT threads; sbuf needs NT/32 floats + NT/32 ints; result valid in all threads
template <int NT>
__device__ __forceinline__ void block_argmin(float& v, int& i, float* sv, int* si) {
    warp_argmin(v, i);
    int w = MBX_TID >> 5, l = MBX_TID & 31;
    MBX_SYNC();
    if (l == 0) { sv[w] = v; si[w] = i; }
    MBX_SYNC();
    v = sv[0]; i = si[0];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) argmin_pair(v, i, sv[k], si[k]);
}
template <int NT>
__device__ __forceinline__ float block_max(float v, float* sv) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = MBX_TID >> 5, l = MBX_TID & 31;
    MBX_SYNC();
    if (l == 0) sv[w] = v;
    MBX_SYNC();
    v = sv[0];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) v = fmaxf(v, sv[k]);
    return v;
}
