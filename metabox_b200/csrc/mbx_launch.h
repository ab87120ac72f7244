// mbx_launch.h -- host-side glue shared by the .cu translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/metabox_b200.h"

void mbx_set_error(const char* fmt, ...);

#define MBX_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            mbx_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return MBX_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

// "done once" flag PER DEVICE: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only, and one process may
// drive several GPUs (InstancePack / Tester / Trainers take a device).  Racing host threads at worst set the attribute twice.
#define MBX_MAX_DEVICES 64
struct MbxPerDevice {
    volatile unsigned char flag[MBX_MAX_DEVICES];
    static int dev() { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < MBX_MAX_DEVICES) ? d : -1; }
    bool get() const { const int d = dev(); return d >= 0 && flag[d]; }
    void set() { const int d = dev(); if (d >= 0) flag[d] = 1; }
};

// BBOB dimensions compiled into the library.  Every kernel is specialised on the dimension (unrolled rotations, shared-memory plans, the
// packed FDR scan over dimension pairs), so the reference's free `--dim` (src/config.py:10) becomes a build-time list: the product build
// carries 10, 20 and 30 (the dimensions the reference's experiments and BASELINE.json use); `python -m metabox_b200.build -DMBX_DIM_EXTRA=<d>`
// adds one more even d <= 32.  M(d, ...) is expanded once per dimension.
#ifdef MBX_DIM_EXTRA
#define MBX_EXTRA_DIM_(M, ...) M(MBX_DIM_EXTRA, __VA_ARGS__)
#else
#define MBX_EXTRA_DIM_(M, ...)
#endif
#define MBX_FOR_BBOB_DIMS(M, ...) M(10, __VA_ARGS__) M(20, __VA_ARGS__) M(30, __VA_ARGS__) MBX_EXTRA_DIM_(M, __VA_ARGS__)
#define MBX_CASE_BBOB_(d, FN, ...) if (cfg->dim == d && kind == MBX_KIND_BBOB) return FN<d, MBX_KIND_BBOB>(__VA_ARGS__);

int mbx_launch_eval(int kind, int D, int x_is_f64, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                    const void* x, void* out, int B, int NP, int subtract_opt, uint64_t seed, uint32_t gen,
                    const double* noise_tape, cudaStream_t st);
int mbx_launch_rs(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                  const double* tape_u, int n_tape_iter, float* curve, int32_t* curve_len, int32_t* fes, int B, cudaStream_t st);
int mbx_launch_rlepso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlepsoState* s, const MbxRlepsoTape* tape, float* feature, int B,
                           cudaStream_t st);
int mbx_launch_rlepso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlepsoState* s, const MbxRlepsoTape* tape, const double* pci,
                           const float* action, float* feature, float* reward, int32_t* dbg, int B, cudaStream_t st);
int mbx_launch_lde_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                        const MbxRunConfig* cfg, const MbxLdeState* s, const MbxLdeTape* tape, float* feature, int B,
                        cudaStream_t st);
int mbx_launch_lde_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                        const MbxRunConfig* cfg, const MbxLdeState* s, const MbxLdeTape* tape, const float* action,
                        float* feature, float* reward, float* dbg, int B, cudaStream_t st);
int mbx_launch_deddqn_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxDdqnState* s, const MbxDdqnTape* tape, float* feature, int B,
                           cudaStream_t st);
int mbx_launch_deddqn_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxDdqnState* s, const MbxDdqnTape* tape, const int32_t* action,
                           float* feature, float* reward, float* dbg, int B, cudaStream_t st);
int mbx_launch_deddqn_run(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                          const MbxDdqnState* s, float* feature, const float* weights, float* ret, int n_steps, int B, cudaStream_t st);
int mbx_launch_gleet_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxGleetState* s, const MbxGleetTape* tape, float* state, int B,
                          cudaStream_t st);
int mbx_launch_gleet_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxGleetState* s, const MbxGleetTape* tape, const float* action,
                          float* state, float* reward, int B, cudaStream_t st);
int mbx_launch_landscape(const double* sample, const double* cost, const double* ref, int B, int S, int D, double* out,
                         cudaStream_t st);
int mbx_launch_qlpso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxQlpsoState* s, const MbxQlpsoTape* tape, int32_t* state, int B,
                          cudaStream_t st);
int mbx_launch_qlpso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                          const MbxRunConfig* cfg, const MbxQlpsoState* s, const MbxQlpsoTape* tape, const int32_t* action,
                          int32_t* state, float* reward, int B, cudaStream_t st);
int mbx_launch_rl_pso_init(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlPsoState* s, const MbxRlPsoTape* tape, float* state, int B,
                           cudaStream_t st);
int mbx_launch_rl_pso_step(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst,
                           const MbxRunConfig* cfg, const MbxRlPsoState* s, const MbxRlPsoTape* tape, const float* action,
                           float* state, float* reward, int B, cudaStream_t st);
int mbx_launch_gleet_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                            float* logits, int B, int NP, cudaStream_t st);
int mbx_launch_deddqn_policy(const float* state, const float* weights, int32_t* action, float* q, int B, cudaStream_t st);
int mbx_launch_rlepso_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                             int B, cudaStream_t st);
int mbx_launch_rl_pso_policy(const float* state, const float* weights, const float* noise, float* action, float* mu, float* sigma,
                             int B, int n_in, cudaStream_t st);
int mbx_launch_qlpso_policy(const int32_t* q_state, const double* q_table, const float* u, int32_t* action, int B, cudaStream_t st);
int mbx_launch_progressive_walk(const double* zone_u, const double* r_u, const int32_t* rD, const double* step_u, double lb, double ub,
                                double step_size, int B, int D, int n_steps, double* out, float* out32, cudaStream_t st);
