// ubench_pipes.cu -- issue rate of the SASS instructions the FDR scan of the RLEPSO step is made of, measured on the GPU it runs on.
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_pipes scripts/ubench_pipes.cu && scripts/ubench_pipes
//
// One CTA of 512 threads on one SM (4 warps per scheduler), every thread runs 8 independent dependency chains of ONE instruction kind for
// ITER iterations (inline PTX, `asm volatile`, so that nothing is folded away); the time is clock64() around the loop (max over warps).
// Printed: warp instructions per cycle and scheduler (SMSP).  1.0 = the issue limit; 0.5 = a pipe that takes one warp instruction every
// two cycles.  "mix" rows interleave two kinds to show which pairs share a pipe (sum of the rates stays at the single-kind rate) and
// which do not (the sum approaches the issue limit).  Measurement helper: not part of the product library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITER 4096
#define NCH 8

#define CHAINS_F(OP)  _Pragma("unroll") for (int c = 0; c < NCH; ++c) { OP(c) }

struct Res { long long cyc; };

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed, int iseed) {
    float x[NCH], y[NCH];
    int u[NCH];
    unsigned long long xx[NCH], yy[NCH];
    double dd[NCH];
    float dd2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        x[c] = seed + c + threadIdx.x * 1e-3f; y[c] = seed * 0.5f + c; u[c] = iseed + c + threadIdx.x;
        xx[c] = ((unsigned long long)__float_as_uint(x[c]) << 32) | __float_as_uint(y[c]); yy[c] = xx[c] ^ 0x0000100000001000ull;
        dd[c] = (double)x[c]; dd2[c] = x[c] * 0.75f;
    }
    const float b = seed * 0.25f + 1.f, one = seed * 0.f + 1.0000001f;
    const int ib = iseed | 1;
    const double db = (double)one;
    int pin = (threadIdx.x & 1) ^ (iseed & 1);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (KIND == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                       // FADD reg
            if (KIND == 1) asm volatile("add.f32 %0, %0, 0f3727C5AC;" : "+f"(x[c]));                         // FADD imm
            if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(one), "f"(b));      // FFMA reg
            if (KIND == 3) asm volatile("fma.rn.f32 %0, %0, %1, 0f3727C5AC;" : "+f"(x[c]) : "f"(one));      // FFMA imm
            if (KIND == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xx[c]) : "l"(yy[c]));             // FADD2
            if (KIND == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx[c]) : "l"(yy[c]), "l"(yy[(c + 1) % NCH]));  // FFMA2
            if (KIND == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(xx[c]) : "l"(yy[c]));             // FMUL2
            if (KIND == 7) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));   // FSEL (+ the predicate is hoisted)
            if (KIND == 8) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y[c]));                     // FMNMX
            if (KIND == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(u[c]) : "r"(ib), "r"(iseed)); // LOP3
            if (KIND == 10) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));                         // MUFU.RCP
            if (KIND == 11) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(ib), "r"(iseed));   // IMAD
            if (KIND == 12) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[c]) : "r"(ib));                      // IADD3
            if (KIND == 13) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dd[c]) : "d"(db));              // DFMA
            if (KIND == 14) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p add.f32 %0, %2, 0f80000000;}" : "+f"(x[c]) : "f"(b), "f"(y[c]));   // FSETP + predicated FADD
            if (KIND == 15) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(x[c]) : "f"(b), "f"(y[c]));          // FSETP + FSEL
            if (KIND == 16) {                                                                                 // mix: FSEL + FFMA2
                asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx[c]) : "l"(yy[c]), "l"(yy[(c + 1) % NCH]));
            }
            if (KIND == 17) {                                                                                 // mix: FSEL + FMNMX (same pipe?)
                asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(y[c]) : "f"(b));
            }
            if (KIND == 18) {                                                                                 // mix: FADD imm + FFMA2
                asm volatile("add.f32 %0, %0, 0f3727C5AC;" : "+f"(x[c]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx[c]) : "l"(yy[c]), "l"(yy[(c + 1) % NCH]));
            }
            if (KIND == 19) {                                                                                 // mix: MUFU.RCP + FFMA2 + FMNMX
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx[c]) : "l"(yy[c]), "l"(yy[(c + 1) % NCH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(y[c]) : "f"(b));
            }
            if (KIND == 20) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %3; selp.f32 %0, %1, %0, p; selp.f32 %1, %3, %1, p; selp.s32 %2, %4, %2, p;}"
                                         : "+f"(x[c]), "+f"(y[c]), "+r"(u[c]) : "f"(b), "r"(ib));                       // the FDR update as it is: FSETP + 3 selects
            if (KIND == 21) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(y[c]), "f"(b));      // FMNMX3
            if (KIND == 22) asm volatile("{.reg .f32 t; abs.f32 t, %0; add.f32 %0, t, 0f3727C5AC;}" : "+f"(x[c]));   // FADD |x|, imm
            if (KIND == 26) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %3; @p sub.f32 %0, 0f80000000, %1; @p add.f32 %1, %3, 0f80000000; selp.s32 %2, %4, %2, p;}"
                                         : "+f"(x[c]), "+f"(y[c]), "+r"(u[c]) : "f"(b), "r"(ib));                       // the FDR update, variant 3: FSETP + 2 predicated FADD + SEL
            if (KIND == 23) asm volatile("sub.f32 %0, %1, %0;" : "+f"(x[c]) : "f"(b));                       // FADD reg, negated source
            if (KIND == 24) asm volatile("mul.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(one));                     // FMUL reg
            if (KIND == 25) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p mov.b32 %0, %2;}" : "+f"(x[c]) : "f"(b), "f"(y[c]));   // FSETP + predicated MOV
            if (KIND == 30) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + FADD reg
                              asm volatile("add.f32 %0, %0, %1;" : "+f"(y[c]) : "f"(b)); }
            if (KIND == 31) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + FFMA reg
                              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[c]) : "f"(one), "f"(b)); }
            if (KIND == 32) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + FADD2
                              asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xx[c]) : "l"(yy[c])); }
            if (KIND == 33) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + IMAD
                              asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(ib), "r"(iseed)); }
            if (KIND == 34) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + 2 FADD
                              asm volatile("add.f32 %0, %0, %1;" : "+f"(y[c]) : "f"(b));
                              asm volatile("add.f32 %0, %0, %1;" : "+f"(dd2[c]) : "f"(b)); }
            if (KIND == 35) { asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(x[c]) : "f"(y[c]), "r"(pin));      // FSEL + DFMA
                              asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dd[c]) : "d"(db)); }
            if (KIND == 36) { asm volatile("min.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                                                             // FMNMX + MUFU
                              asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(y[c])); }
            if (KIND == 37) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                                                             // FADD + MUFU
                              asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(y[c])); }
            if (KIND == 38) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                                                             // FADD + DFMA
                              asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dd[c]) : "d"(db)); }
            if (KIND == 39) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                                                             // FADD + FMUL (both FMA pipe)
                              asm volatile("mul.f32 %0, %0, %1;" : "+f"(y[c]) : "f"(one)); }
            if (KIND == 40) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(b));                                                             // FADD + IADD3
                              asm volatile("add.s32 %0, %0, %1;" : "+r"(u[c]) : "r"(ib)); }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc += x[c] + y[c] + dd2[c] + (float)u[c] + (float)(xx[c] & 0xffff) + (float)dd[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + b;
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

template <int KIND>
static void run(const char* name, int ops_per_chain_iter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 512 * sizeof(float)); cudaMalloc(&cyc, 16 * sizeof(long long));
    k<KIND><<<1, 512>>>(out, cyc, 1.25f, 3);
    k<KIND><<<1, 512>>>(out, cyc, 1.25f, 3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 16; ++i) mx = h[i] > mx ? h[i] : mx;
    const double winst = 4.0 * ITER * NCH * ops_per_chain_iter;        // warp instructions per scheduler (4 warps each)
    printf("%-34s %8.3f warp-instructions / cycle / scheduler   (%lld cycles)\n", name, winst / (double)mx, mx);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("# %s, %d SMs; one CTA of 512 threads, %d chains per thread, %d iterations\n", p.name, p.multiProcessorCount, NCH, ITER);
    run<0>("FADD  reg", 1);            run<1>("FADD  imm", 1);          run<23>("FADD  reg, -src", 1);   run<24>("FMUL  reg", 1);
    run<2>("FFMA  reg", 1);            run<3>("FFMA  imm", 1);
    run<4>("FADD2", 1);                run<5>("FFMA2", 1);              run<6>("FMUL2", 1);
    run<7>("FSEL", 1);                 run<8>("FMNMX", 1);              run<21>("FMNMX3", 1);            run<9>("LOP3", 1);
    run<12>("IADD3", 1);               run<11>("IMAD", 1);              run<10>("MUFU.RCP", 1);          run<13>("DFMA", 1);
    run<15>("FSETP + FSEL (2)", 2);    run<14>("FSETP + @p FADD (2)", 2);  run<25>("FSETP + @p MOV (2)", 2);
    run<20>("FSETP + 3 selects (4)", 4);   run<26>("FSETP + 2 @p FADD + SEL (4)", 4);   run<22>("FADD |x|, imm", 1);
    run<16>("mix FSEL + FFMA2 (2)", 2);   run<17>("mix FSEL + FMNMX (2)", 2);   run<18>("mix FADD imm + FFMA2 (2)", 2);
    run<19>("mix MUFU.RCP + FFMA2 + FMNMX (3)", 3);
    run<30>("mix FSEL + FADD (2)", 2);   run<34>("mix FSEL + 2 FADD (3)", 3);   run<31>("mix FSEL + FFMA (2)", 2);   run<32>("mix FSEL + FADD2 (2)", 2);
    run<33>("mix FSEL + IMAD (2)", 2);   run<35>("mix FSEL + DFMA (2)", 2);     run<36>("mix FMNMX + MUFU.RCP (2)", 2);   run<37>("mix FADD + MUFU.RCP (2)", 2);
    run<38>("mix FADD + DFMA (2)", 2);   run<39>("mix FADD + FMUL (2)", 2);     run<40>("mix FADD + IADD3 (2)", 2);
    return 0;
}
