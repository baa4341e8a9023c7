// Shared device helpers for libpmce_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PMCE_WARP 32

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

// exact-erf GELU (nn.GELU default; reference timm Mlp act_layer=nn.GELU)
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
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
