// mbx_hostdev.cuh -- the few helpers shared by the policy headers that compile BOTH for the device (nvcc) and for the host
// emulation of the CPU tests (g++, tests/emul/): 16-byte vector loads / stores that degrade to scalar accesses on the host.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MBX_HD __host__ __device__ __forceinline__
#else
#define MBX_HD inline
#endif

namespace mbxhd {
struct F4 { float x, y, z, w; };
MBX_HD F4 ld4(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = *reinterpret_cast<const float4*>(p);
    return F4{v.x, v.y, v.z, v.w};
#else
    return F4{p[0], p[1], p[2], p[3]};
#endif
}
MBX_HD void st4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}
}  // namespace mbxhd
