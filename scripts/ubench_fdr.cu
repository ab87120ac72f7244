// ubench_fdr.cu -- the FDR screening scan of the RLEPSO step (metabox_b200/csrc/mbx_fdr.cuh, the SAME code the product kernel compiles) alone,
// at the occupancy it has inside the step kernel (128 threads, 4 CTAs per SM, 56 KB of shared memory per CTA), on synthetic swarms:
// the A/B bench for formulations of the inner loop (compile-time knobs, -D...), which the full bench resolves only to +-0.5 %.
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Imetabox_b200/csrc [-DMBX_FDR_PRED=3 ...] \
//          -o scripts/ubench_fdr_<name> scripts/ubench_fdr.cu
//
// Prints the time of one scan of a 100 x 30 swarm per CTA, the cycles per 8 (pair, dimension) comparisons and scheduler, the share of
// ambiguous items and a checksum of the unambiguous targets (must be the same for every variant).  Measurement helper, not product code.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "mbx_fdr.cuh"

#ifndef VARIANT
#define VARIANT "default"
#endif
#ifndef UB_CTAS
#define UB_CTAS 4                   // CTAs per SM (4 = the step kernel; more = what the scan would do with more warps per scheduler)
#endif
constexpr int D = 30, NP = 100, REP = 200;

__global__ void __launch_bounds__(128, UB_CTAS)
fdr_bench(const float* __restrict__ gx, const float* __restrict__ gf, unsigned long long* sums, int* namb) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* spx = reinterpret_cast<float*>(smem);                                  // [NP*D]
    float2* spf2 = reinterpret_cast<float2*>(spx + NP * D);                       // [NP]
    unsigned char* order = reinterpret_cast<unsigned char*>(spf2 + NP);           // [128]
    unsigned char* tgt = order + 128;                                             // [NP*D]
    unsigned short* fbl = reinterpret_cast<unsigned short*>(tgt + ((NP * D + 15) & ~15));   // [NP*D]
    int* nfb = reinterpret_cast<int*>(fbl + NP * D);
    const int tid = threadIdx.x, b = blockIdx.x;
    for (int e = tid; e < NP * D; e += 128) spx[e] = gx[(size_t)(b % 64) * NP * D + e];
    if (tid < NP) { const float f = gf[(b % 64) * NP + tid]; fdr_store_cost(spf2, tid, f); order[tid] = (unsigned char)tid; }
    unsigned long long acc = 0; int amb = 0;
    for (int r = 0; r < REP; ++r) {
        if (tid == 0) *nfb = 0;
        __syncthreads();
        fdr_screen<D>(NP, tid, spx, spf2, order, tgt, fbl, nfb);
        __syncthreads();
        if (r == REP - 1) {
            // checksum over the items that were NOT flagged (flagged ones are settled by the float64 pass of the real kernel)
            for (int e = tid; e < NP * D; e += 128) {
                bool flagged = false;
                for (int k = 0; k < *nfb; ++k) flagged |= (fbl[k] == e);
                if (!flagged) acc += (unsigned long long)(tgt[e] + 1) * (e + 1);
            }
            amb = *nfb;
        }
        __syncthreads();
    }
    atomicAdd(sums, acc);
    if (tid == 0) atomicAdd(namb, amb);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsw = 64;
    std::vector<float> hx((size_t)nsw * NP * D), hf((size_t)nsw * NP);
    unsigned long long st = 88172645463325252ull;
    auto u01 = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    for (auto& v : hx) v = (float)(-5. + 10. * u01());
    for (int s = 0; s < nsw; ++s) {
        for (int i = 0; i < NP; ++i) hf[s * NP + i] = (float)(1e3 * u01() * u01());
        std::sort(hf.begin() + s * NP, hf.begin() + (s + 1) * NP);
        if (s % 4 == 1) for (int e = 0; e < NP * D; e += 7) hx[(size_t)s * NP * D + e] = 5.f;     // particles sitting on the bound: equal distances
        if (s % 4 == 2) for (int i = 10; i < 20; ++i) hf[s * NP + i] = hf[s * NP + 10];           // a plateau of equal costs
    }
    float *gx, *gf; unsigned long long* sums; int* namb;
    cudaMalloc(&gx, hx.size() * 4); cudaMalloc(&gf, hf.size() * 4); cudaMalloc(&sums, 8); cudaMalloc(&namb, 4);
    cudaMemcpy(gx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(gf, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (UB_CTAS == 4 ? 56 : 224 / UB_CTAS - 1) * 1024;
    cudaFuncSetAttribute(fdr_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fdr_bench, 128, smem);
    const int grid = p.multiProcessorCount * UB_CTAS;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaMemset(sums, 0, 8); cudaMemset(namb, 0, 4);
        cudaEventRecord(e0);
        fdr_bench<<<grid, 128, smem>>>(gx, gf, sums, namb);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    unsigned long long hs; int ha; cudaMemcpy(&hs, sums, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&ha, namb, 4, cudaMemcpyDeviceToHost);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cmp = (double)NP * (NP - 1) / 2 * D;                       // (pair, dimension) comparisons per scan
    const double us_scan = 1e3 * best / REP;                                // 4 CTAs per SM scan concurrently
    // per scheduler: 4 CTAs x 4 warps on 4 schedulers -> each scheduler serves 4 warps = one CTA's worth of comparisons per scan time
    const double cyc8 = us_scan * 1e-6 * clk * 1e3 / (cmp / 32.0 / 8.0) * 4.0 / UB_CTAS;
    printf("variant %-28s occupancy %d CTAs/SM  %8.2f us per scan (that many per SM in flight)  %6.1f cycles per 8 comparisons and scheduler (at %d MHz nominal)  "
           "ambiguous %.3f %%  checksum %llu\n", VARIANT, occ, us_scan, cyc8, clk / 1000, 100.0 * ha / ((double)grid * NP * D), hs);
    return 0;
}
