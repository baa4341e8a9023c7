// Shared device helpers for libpmce_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PMCE_WARP 32

// Programmatic dependent launch (PDL): every kernel of the forward is launched with programmatic stream serialisation
// (host_once.h pmce_launch), so its CTAs may be scheduled - and run their prologue: barrier init, TMEM allocation, descriptor
// prefetch - while the previous kernel of the stream is still draining. pdl_wait() blocks until that kernel has COMPLETED and
// its writes are visible; nothing before it may touch global memory another kernel writes or reads-then-overwrites.
// EVERY kernel launched through pmce_launch must execute pdl_wait() (completion is transitive only through it), and kernels
// that allocate TMEM call pdl_trigger() only after the allocation (a dependent CTA that lands on the same SM and takes the
// columns first would otherwise block this CTA's allocation for ever). Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }     // kernels without a prologue worth overlapping

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// erf-form GELU (nn.GELU default; reference timm Mlp act_layer=nn.GELU): 0.5 x (1 + erf(x / sqrt 2)).
// erf by Abramowitz & Stegun 7.1.26 (|err| <= 1.5e-7, the size of fp32 rounding): one rcp, one ex2 and a degree-5 Horner
// instead of erff's two-branch polynomial — about half the instructions of the fc1 epilogue, which is issue-bound.
// Measured against float64 over [-8, 8]: max |gelu error| 4.7e-7 (fp32 erff-based: 4.5e-7).
__device__ __forceinline__ float erf_as(float z) {
    const float az = fabsf(z);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    p *= t;
    const float r = fmaf(-p, __expf(-az * az), 1.0f);
    return copysignf(r, z);
}
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Generalised row -> element-offset mapping used by GEMM outputs and attention row addressing:
//   off(r) = (r / div) * s0 + (r % div) * s1
struct RowMap {
    int div;
    long long s0, s1;
    __host__ __device__ __forceinline__ long long operator()(int r) const {
        return (long long)(r / div) * s0 + (long long)(r % div) * s1;
    }
};
