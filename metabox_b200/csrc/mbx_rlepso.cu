// mbx_rlepso.cu -- fused RLEPSO generation (PSO family): velocity components (CLPSO tournament, FDR
// argmin, gbest, pbest) -> clip -> position -> fitness -> pbest/gbest -> patience -> stochastic
// re-initialisation -> cost-curve log, one CTA per lane, state float32 in HBM / float64 in registers.
// Reference: optimizer/rlepso_optimizer.py:39-65 (init_population), :76-132 (velocity parts,
// coefficients), :134-168 (re-init), :173-263 (update).  numpy twin: oracle/rlepso.py.
#include <stdlib.h>
#include "mbx_eval.cuh"
#include "mbx_launch.h"
#include "mbx_fdr.cuh"


#ifdef MBX_PROFILE
extern "C" int mbx_prof_read(unsigned long long* h16, int reset) {
    if (cudaMemcpyFromSymbol(h16, g_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return MBX_ECUDA;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return MBX_ECUDA;
    }
    return MBX_OK;
}
#endif

// Shared-memory plan, sized by the run-time population size (NP = 100, D = 30 -> 56.0 KB: 4 CTAs per SM).
// Lifetimes: region A holds the sorted pbest_x (f32) and the FDR targets during P0-P2 and the evaluation scratch in P3;
// the block [spf2 .. order] is dead during P3 and doubles as the Gallagher peak-table staging tile.
struct RlLayout {
    int NP, NPp;
    size_t oConst, oA, oSpf, oR1, oR2, oGco, oRank, oOrder, oTileEnd, oGx, oTmpf, oUr, oXs, oF, oPni, oMask, oRed, bytes;
    __host__ __device__ RlLayout(int np, int D, int kConst, bool protein) {
        NP = np; NPp = (np + 3) & ~3;
        const size_t scratch = protein ? (size_t)(MBX_NT / 32) * MBX_PROT_WARP_SCRATCH : (size_t)(D + 1) * NPp;   // bbob_eval_cta tile [NP][D+1]
        size_t regionA = scratch > (size_t)NPp * D ? scratch : (size_t)NPp * D;
        if (regionA < MBX_NT) regionA = MBX_NT;              // >= 128 x 8 bytes: exchange buffer of the sorting network (block_sort_order_128)
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oA = oConst + sizeof(double) * kConst;               // spx f32 [NP*D] | tgt u8 [NP*D]   /   eval scratch f64
        oSpf = oA + sizeof(double) * regionA;                // sorted pbest duplicated (f, f) float2 [NP]
        oR1 = oSpf + sizeof(float2) * NPp;                   // rand1 as Philox word u32 [NP]
        oR2 = oR1 + 4 * NPp;                                 // rand2
        oGco = oR2 + 4 * NPp;                                // 5 groups x (w, c1..c4) f64
        oRank = oGco + sizeof(double) * 26;                  // rank / order u8 [NP]
        oOrder = oRank + NPp;
        oTileEnd = oOrder + NPp;                             // --- end of the block that is dead during P3 ---
        oGx = (oTileEnd + 7) & ~(size_t)7;                   // gbest_x as double [D]
        oTmpf = oGx + sizeof(double) * D;                    // re-initialisation costs f32 [NP]
        oUr = oTmpf + 4 * NPp;                               // re-init uniform as Philox word u32 [NP]
        oXs = (oUr + 4 * NPp + 15) & ~(size_t)15;            // new positions f32 [NP*D]  (P1: list of ambiguous FDR items)
        oF = oXs + sizeof(float) * NPp * D;                  // pf32, nc, cc f32 [NP]
        oPni = oF + sizeof(float) * 3 * NPp;                 // per_no_improve i32 [NP]
        oMask = oPni + 4 * NPp;                              // re-init mask, better flags u8 [NP]
        oRed = (oMask + 2 * NPp + 15) & ~(size_t)15;         // reduction scratch + a0 coefficients + list counter
        bytes = oRed + 96;
    }
};

template <int D, int KIND>
struct RlepsoViews {
    MbxInstance* h;
    double *sC, *A, *gco, *gx;
    float2* spf2;
    uint32_t *r1u, *r2u, *uru;
    float *spx, *xs, *pf32, *nc, *cc, *tmpf, *a0c;
    int* pni;
    unsigned char *rank, *order, *tgt, *mask, *better, *tile;
    int tile_bytes;
    unsigned short* fbl;        // FDR items that need the exact float64 pass (aliases xs: written before P2 fills xs)
    int* nfb;
    float* redv;
    int* redi;
    __device__ __forceinline__ RlepsoViews(unsigned char* s, const RlLayout& L) {
        const int N = L.NPp;
        h = reinterpret_cast<MbxInstance*>(s);
        sC = reinterpret_cast<double*>(s + L.oConst);
        A = reinterpret_cast<double*>(s + L.oA);
        spx = reinterpret_cast<float*>(s + L.oA);
        tgt = s + L.oA + sizeof(float) * N * D;
        spf2 = reinterpret_cast<float2*>(s + L.oSpf);
        r1u = reinterpret_cast<uint32_t*>(s + L.oR1);
        r2u = reinterpret_cast<uint32_t*>(s + L.oR2);
        uru = reinterpret_cast<uint32_t*>(s + L.oUr);
        gco = reinterpret_cast<double*>(s + L.oGco);
        gx = reinterpret_cast<double*>(s + L.oGx);
        rank = s + L.oRank; order = s + L.oOrder;
        tmpf = reinterpret_cast<float*>(s + L.oTmpf);
        tile = s + L.oSpf; tile_bytes = (int)(L.oTileEnd - L.oSpf);
        xs = reinterpret_cast<float*>(s + L.oXs);
        float* f = reinterpret_cast<float*>(s + L.oF);
        pf32 = f; nc = f + N; cc = f + 2 * N;
        pni = reinterpret_cast<int*>(s + L.oPni);
        mask = s + L.oMask; better = mask + N;
        redv = reinterpret_cast<float*>(s + L.oRed);
        redi = reinterpret_cast<int*>(s + L.oRed + 32);
        nfb = reinterpret_cast<int*>(s + L.oRed + 56);
        a0c = reinterpret_cast<float*>(s + L.oRed + 64);
        fbl = reinterpret_cast<unsigned short*>(xs);
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

// Compile-time tuning knobs (defaults = the measured configuration; scripts/variant_sweep.sh builds and benches alternatives):
#ifndef MBX_RLEPSO_MIN_CTAS
#define MBX_RLEPSO_MIN_CTAS 4       // __launch_bounds__ minimum CTAs per SM of the step kernel (4 -> 128 registers)
#endif
#ifndef MBX_RLEPSO_P2_CH
#define MBX_RLEPSO_P2_CH 12         // elements per thread whose x / v loads are issued ahead in the velocity phase
#endif
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
                                         MBX_TAG_NOISE, cfg.use_tape && T.noise ? T.noise + (size_t)b * 6 * NP : nullptr, NP, nullptr, L.NPp,
                                         V.tile, V.tile_bytes);   // P0-P2 per-particle scratch is dead here: Gallagher peak-table staging tile
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
// TEST = true is the instantiation the parity tests drive: draws may come from a tape of the oracle's float64 uniforms (cfg.use_tape) and
// the FDR targets can be written out (dbg).  TEST = false, what every sweep launches, contains neither: the tape's loads, 64-bit address
// arithmetic and branches were issued (predicated off) for every element of the product path before.  Same source otherwise; the device
// tests compare the two instantiations bit for bit on Philox draws (tests/test_gpu_rlepso.py).
template <int D, int KIND, bool TEST>
__global__ void __launch_bounds__(MBX_NT, MBX_RLEPSO_MIN_CTAS)
rlepso_step_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                   MbxRunConfig cfg, MbxRlepsoState S, MbxRlepsoTape T, const double* __restrict__ pci,
                   const float* __restrict__ action, float* __restrict__ feature, float* __restrict__ reward,
                   int32_t* __restrict__ dbg) {
    const int b = blockIdx.x;
    if (S.done[b]) { if (threadIdx.x == 0) reward[b] = 0.f; return; }      // finished lane: state frozen, reward 0
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PROF_INIT
    const int NP = cfg.NP, tid = threadIdx.x;
    const RlLayout L(NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN);
    RlepsoViews<D, KIND> V(smem_raw, L);
    const MbxInstance* gi = inst + lane_inst[b];
    const bool use_tape = TEST && cfg.use_tape;
    const size_t base = (size_t)b * NP * D, pb = (size_t)b * NP;
    const uint32_t gen = (uint32_t)S.gen[b];
    Philox ph(cfg.seed);
    // the pbest_x rows are requested NOW (NP * D / 128 <= D values per thread, NP <= 128) and scattered into their rank order after P0b:
    // their DRAM latency hides behind the constant staging, P0 and the rank loop instead of being exposed between two barriers
    float pbx[D];
    {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int e = tid + k * MBX_NT;
            pbx[k] = e < NP * D ? __ldcs(S.pbest_x + base + e) : 0.f;
        }
    }
    stage_instance<D, KIND>(gi, consts, V.h, V.sC);          // (after the row requests: its stores wait for its own loads)
    // ---- P0: per-particle scalars, draws, coefficients (:112-132) -------------------------------------------
    for (int d = tid; d < D; d += MBX_NT) V.gx[d] = (double)S.gbest_x[(size_t)b * D + d];
    if (tid < NP) {
        const int i = tid;
        V.pf32[i] = S.pbest[pb + i];
        V.cc[i] = S.c_cost[pb + i]; V.pni[i] = S.pni[pb + i];
        if (!use_tape) {                                          // tape mode reads its float64 draws from the tape directly
            uint4 w4 = ph((uint32_t)i, gen, (uint32_t)(cfg.lane_base + b), MBX_TAG_P);
            V.r1u[i] = w4.x; V.r2u[i] = w4.y; V.uru[i] = w4.z;
        }
    }
    if (tid < 5) {                                                   // float32 coefficient math (numpy-2 promotion)
        const float* a = action + (size_t)b * 35 + tid * 5;
        {
            const float a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6];
            const float w = a1 * 0.8f + 0.1f;
            float ssum = a3 + a4; ssum = ssum + a5; ssum = ssum + a6; ssum = ssum + 1e-5f;
            float scale = 1.f / ssum; scale = scale * a2; scale = scale * 8.f;
            double* gc = V.gco + tid * 5;
            gc[0] = (double)w; gc[1] = (double)(scale * a3); gc[2] = (double)(scale * a4);
            gc[3] = (double)(scale * a5); gc[4] = (double)(scale * a6);
        }
        V.a0c[tid] = a[0] * 0.01f;                 // c_mutation = float32(a0 * 0.01) * per_no_improve (:119)
    }
    if (tid < NP) {                                                  // coefficient group of every particle, 20 % of the swarm each (:113-118): the
        const int per = NP / 5;                                       // run-time division happens once per particle here, not once per (i, d) in P2
        V.mask[tid] = (unsigned char)min(tid / per, 4);
    }
    __syncthreads();
    PROF_MARK(0);
    const double lb = V.h->lb, ub = V.h->ub, vmax = 0.1 * (ub - lb);
    {
    // ---- P0b: stable rank of the pbest costs; sorted float32 copies of pbest / pbest_x (both ARE float32 in HBM) ----
    float* spx = V.spx;                                  // pbest_x rows in ascending-pbest order [NP*D]
    float2* spf2 = V.spf2;                               // sorted pbest costs (layout: mbx_fdr.cuh, FDR_COST)
    {
        // 128-thread sorting network on (order key, index) items (mbx_common.cuh); region A is written after this: exchange buffer
        const int at = block_sort_order_128(tid < NP ? V.pf32[tid] : 0.f, NP, reinterpret_cast<unsigned long long*>(V.A));
        if (tid < NP) { V.rank[at] = (unsigned char)tid; V.order[tid] = (unsigned char)at; fdr_store_cost(spf2, tid, V.pf32[at]); }
    }
    if (tid == 0) *V.nfb = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int e = tid + k * MBX_NT;
        if (e < NP * D) { const int i = e / D; spx[V.rank[i] * D + (e - i * D)] = pbx[k]; }
    }
    __syncthreads();
    PROF_MARK(1);
    // ---- P1: FDR targets  argmin_j (pf[j]-pf[i]) / (|px[j,d]-px[i,d]|+1e-5), first minimum   (:97-103) ----
    // The minimum is <= 0 (j = i gives 0), so only particles with a strictly better pbest -- the ones BEFORE i in the
    // sorted order -- can win (equal costs tie with j = i at ratio 0 and are resolved by the index rule).
    // Pass 1 (all items): float32 screening (mbx_fdr.cuh: quotient keys a * rcp(q) carrying the candidate's position, smallest and second
    // smallest key per item); an item whose two smallest keys are closer than the rounding bound, or whose costs leave the range the bound
    // covers, is marked ambiguous.  Pass 2 (ambiguous items only, compacted): the reference's own float64 divisions and first-index argmin.
    // An unambiguous float32 winner beats every other candidate by more than the error of both keys, so the float64 quotients are ordered
    // the same way and the result equals np.argmin exactly.
    fdr_screen<D>(NP, tid, spx, spf2, V.order, V.tgt, V.fbl, V.nfb);
    __syncthreads();
    {
        // Pass 2, eight lanes per ambiguous item: the float32 winner js sits within 3.3e-5 (two key errors) of the
        // true minimum, so only candidates whose float32 quotient is within FDR_NEAR of it can be the float64 argmin; those
        // (a handful) get the reference's float64 division, and the (quotient, original index) minimum is reduced by shuffles.
        const int nfb = *V.nfb, grp = tid >> 3, gl = tid & 7;
        PROF_ADD(6, nfb); PROF_ADD(7, NP * D);
        for (int k0 = 0; k0 < nfb; k0 += MBX_NT / 8) {                 // uniform trip count: shuffles below are warp-wide
            const int k = k0 + grp;
            const bool on = k < nfb;
            const int item = on ? V.fbl[k] : 0, r = item / D, d = item - r * D;
            const int jt = V.tgt[V.order[r] * D + d];                 // float32 winner (sorted position, == r for "self")
            const bool every = jt == FDR_NO_WINNER;                   // no usable winner: the float64 division for every candidate
            const int jb = every ? r : jt;
            const float pxr = spx[r * D + d], pfr = FDR_COST(spf2, r);
            const float ab = FDR_COST(spf2, jb) - pfr, qb = fabsf(spx[jb * D + d] - pxr) + 1e-5f;
            double best = 0.; int bs = r, bo = V.order[r];            // j = i: quotient 0
            for (int s = gl; s < r && on; s += 8) {
                const float a = FDR_COST(spf2, s) - pfr, dq = spx[s * D + d] - pxr, q = fabsf(dq) + 1e-5f;
                const float lhs = a * qb, rhs = ab * q;               // a/q vs ab/qb (both <= 0)
                if (every || !(fabsf(lhs - rhs) > fmaf(fabsf(rhs), FDR_NEAR, FDR_TINY)) || s == jb) {
                    const double a64 = (double)FDR_COST(spf2, s) - (double)pfr;
                    const double q64 = a64 == 0. ? 0. : a64 / (fabs((double)spx[s * D + d] - (double)pxr) + 1e-5);
                    const int o = V.order[s];
                    if (q64 < best || (q64 == best && o < bo)) { best = q64; bs = s; bo = o; }
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int os = __shfl_xor_sync(0xffffffffu, bs, o), oo = __shfl_xor_sync(0xffffffffu, bo, o);
                if (ob < best || (ob == best && oo < bo)) { best = ob; bs = os; bo = oo; }
            }
            if (on && gl == 0) V.tgt[V.order[r] * D + d] = (unsigned char)bs;
        }
    }
    __syncthreads();
    PROF_MARK(2);
    // ---- P2: velocity + position, element-wise over (i,d)   (:76-95, :105-109, :179-195) ---------------
    // the loads of x and v for a chunk of P2_CH elements are issued before their first use: one exposed DRAM latency per
    // chunk instead of one per element
    constexpr int P2_CH = MBX_RLEPSO_P2_CH;
#pragma unroll 1
    for (int k0 = 0; k0 * MBX_NT < NP * D; k0 += P2_CH) {
    float xr[P2_CH], vr[P2_CH];
    // one 64-bit pointer per array and chunk; the elements of the chunk sit at compile-time offsets from it
    float* const xg = S.x + base + tid + k0 * MBX_NT;
    float* const vg = S.v + base + tid + k0 * MBX_NT;
#pragma unroll
    for (int k = 0; k < P2_CH; ++k) {
        const int e = tid + (k0 + k) * MBX_NT;
        xr[k] = 0.f; vr[k] = 0.f;
        if (e < NP * D) { xr[k] = __ldcs(xg + k * MBX_NT); vr[k] = __ldcs(vg + k * MBX_NT); }
    }
#pragma unroll
    for (int k = 0; k < P2_CH; ++k) {
        const int e = tid + (k0 + k) * MBX_NT;
        if (e >= NP * D) break;
        const int i = e / D, d = e - i * D;
        double ru, fu; int t0, t1;
        if (use_tape) {
            ru = T.clpso_u[base + e]; fu = T.fdr_u[base + e];
            t0 = T.tour[(base + e) * 2]; t1 = T.tour[(base + e) * 2 + 1];
        } else {
            uint4 w = ph((uint32_t)e, gen, (uint32_t)(cfg.lane_base + b), MBX_TAG_PD);
            ru = u01(w.x); t0 = bounded(w.y, NP); t1 = bounded(w.z, NP); fu = u01(w.w);
        }
        const double x = (double)xr[k], v = (double)vr[k];
        const double pxi = (double)spx[V.rank[i] * D + d];
        const int tsel = (V.pf32[t1] < V.pf32[t0]) ? t1 : t0;          // argmin of the 2 picks, tie -> first
        const double own = (ru > pci[i]) ? pxi : (double)spx[V.rank[tsel] * D + d];
        const double v_clpso = ru * (own - x);
        const int tg = V.tgt[e];                                       // sorted position of the FDR target
        if (TEST && dbg) dbg[base + e] = V.order[tg];
        const double v_fdr = fu * ((double)spx[tg * D + d] - pxi);
        const double r1 = use_tape ? T.rand1[pb + i] : u01(V.r1u[i]), r2 = use_tape ? T.rand2[pb + i] : u01(V.r2u[i]);
        const double v_pbest = r1 * (pxi - x);
        const double v_gbest = r2 * (V.gx[d] - x);
        const double* gc = V.gco + V.mask[i] * 5;                      // group of particle i (table written in P0; `mask` is free until P4)
        double nv = gc[0] * v + gc[1] * v_clpso + gc[2] * v_fdr + gc[3] * v_gbest + gc[4] * v_pbest;
        // np.clip as two compares and selects each (fmin / fmax add NaN handling the reference does not have: np.clip keeps a NaN)
        nv = nv < -vmax ? -vmax : nv; nv = nv > vmax ? vmax : nv;
        double nx = x + nv;
        nx = nx < lb ? lb : nx; nx = nx > ub ? ub : nx;
        const float nx32 = (float)nx;
        V.xs[e] = nx32;
        xg[k * MBX_NT] = nx32; vg[k * MBX_NT] = (float)nv;
    }
    }
    }
    const MbxInstance& h = *V.h;
    __syncthreads();
    PROF_MARK(3);
    // ---- P3: fitness of the NP new positions (:68-74, :198) ---------------------------------------------
    cta_eval_tile<D, KIND, float, float>(h, V.sC, consts + gi->off, V.xs, NP, 0, V.A, V.nc, true, ph, gen, (uint32_t)(cfg.lane_base + b),
                                         MBX_TAG_NOISE, use_tape && T.noise ? T.noise + (size_t)b * 6 * NP : nullptr, NP, nullptr, L.NPp,
                                         V.tile, V.tile_bytes);   // P0-P2 per-particle scratch is dead here: Gallagher peak-table staging tile
    __syncthreads();
    PROF_MARK(4);
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
        const int pni0 = V.pni[i];
        int pni = (ncv < V.cc[i]) ? 0 : pni0 + 1;
        V.pni[i] = pni;
        V.cc[i] = ncv;
        const int per = NP / 5;
        int g = i / per; if (g > 4) g = 4;
        const double cm = (double)V.a0c[g] * (double)pni0;            // float32 coefficient x per_no_improve of the old state
        const double ur = use_tape ? T.reinit_u[pb + i] : u01(V.uru[i]);
        mymask = ur < cm * 0.01 * (double)pni;
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
            if (use_tape) { up = T.reinit_pos_u[base + e]; uv = T.reinit_vel_u[base + e]; }
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
                                             use_tape && T.noise ? T.noise + (size_t)b * 6 * NP + 3 * NP : nullptr, NP,
                                             noisy ? nullptr : V.mask, L.NPp, V.tile, V.tile_bytes);
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
    PROF_MARK(5);
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
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done.set();
    }
    MbxRlepsoTape t{};
    if (tape) t = *tape;
    kern<<<B, MBX_NT, RlLayout(cfg->NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN).bytes, st>>>(inst, consts, lane_inst, *cfg, *s, t, feature);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
template <int D, int KIND, bool TEST>
static int launch_step_tt(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                          const MbxRlepsoState* s, const MbxRlepsoTape& t, const double* pci, const float* action,
                          float* feature, float* reward, int32_t* dbg, int B, cudaStream_t st) {
    const size_t full = RlLayout(cfg->NP, D, EvalTraits<D, KIND>::kConstDoubles, KIND == MBX_KIND_PROTEIN).bytes;
    auto kern = rlepso_step_kernel<D, KIND, TEST>;
    static MbxPerDevice attr_done;
    if (!attr_done.get()) {
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done.set();
    }
    kern<<<B, MBX_NT, full, st>>>(inst, consts, lane_inst, *cfg, *s, t, pci, action, feature, reward, dbg);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}
// One generation = one launch.  A tape or a debug buffer selects the TEST instantiation (see the kernel).
template <int D, int KIND>
static int launch_step_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                         const MbxRlepsoState* s, const MbxRlepsoTape* tape, const double* pci, const float* action,
                         float* feature, float* reward, int32_t* dbg, int B, cudaStream_t st) {
    MbxRlepsoTape t{};
    if (tape) t = *tape;
    if (cfg->use_tape || dbg) return launch_step_tt<D, KIND, true>(inst, consts, lane_inst, cfg, s, t, pci, action, feature, reward, dbg, B, st);
    return launch_step_tt<D, KIND, false>(inst, consts, lane_inst, cfg, s, t, pci, action, feature, reward, dbg, B, st);
}

#define MBX_DISPATCH(FN, ...)                                                                         \
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, FN, __VA_ARGS__)                                                 \
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
