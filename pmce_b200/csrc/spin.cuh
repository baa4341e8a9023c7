// (f)2, the step BEFORE the path: the SPIN / HMR ResNet-50 feature extractor (reference lib/models/spin.py:129-143, called per
// frame crop at main/run_demo.py:315) that produces the 2048-d image features PMCE.forward consumes.
// Activations are NHWC so every convolution is a GEMM over the channel axis on the tcgen05 GEMM (gemm_tc.cuh, split-bf16):
//   1x1 conv          = [B H W, Cin] x [Cout, Cin]^T                       (the stream itself is the A operand)
//   3x3 / 7x7 conv    = im2col rows [B Ho Wo, kh kw Cin] x [Cout, kh kw Cin]^T   (explicit im2col in split-bf16: 16-byte gathers)
// BatchNorm (eval) is folded into the weights and a bias at pack time (pmce_b200/spin.py); ReLU rides in the GEMM epilogue
// (split-bf16 outputs) or in the small kernels below. The kernels here are the data-movement glue: stem im2col, max-pool,
// 3x3 im2col, strided row gather, ReLU + split, average pool.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"

constexpr int SPIN_STEM_K = 152;          // 7 * 7 * 3 = 147 padded to a multiple of 8

// frames [B,3,224,224] NCHW fp32 -> im2col rows of the 7x7 / stride 2 / pad 3 stem: A[(b,ho,wo), (kh*7+kw)*3+c], split-bf16
__global__ void spin_stem_im2col_kernel(const float* __restrict__ x, int B, SplitOut A) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 K-elements
    const size_t total = (size_t)B * 112 * 112 * (SPIN_STEM_K / 8);
    if (idx >= total) return;
    const int k8 = (int)(idx % (SPIN_STEM_K / 8));
    const size_t m = idx / (SPIN_STEM_K / 8);
    const int wo = (int)(m % 112), ho = (int)((m / 112) % 112), b = (int)(m / (112 * 112));
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = k8 * 8 + i;
        float val = 0.f;
        if (k < 147) {
            const int c = k % 3, kw = (k / 3) % 7, kh = k / 21;
            const int hi = ho * 2 + kh - 3, wi = wo * 2 + kw - 3;
            if (hi >= 0 && hi < 224 && wi >= 0 && wi < 224) val = x[(((size_t)b * 3 + c) * 224 + hi) * 224 + wi];
        }
        v[i] = val;
    }
    uint4 hh, ll;
    tc::split_bf16x2(v[0], v[1], hh.x, ll.x); tc::split_bf16x2(v[2], v[3], hh.y, ll.y);
    tc::split_bf16x2(v[4], v[5], hh.z, ll.z); tc::split_bf16x2(v[6], v[7], hh.w, ll.w);
    *reinterpret_cast<uint4*>(A.hi + m * SPIN_STEM_K + k8 * 8) = hh;
    *reinterpret_cast<uint4*>(A.lo + m * SPIN_STEM_K + k8 * 8) = ll;
}

// ReLU + MaxPool2d(3, stride 2, pad 1) (spin.py:132-134; ReLU and max commute): y [B,112,112,64] fp32 NHWC -> [B,56,56,64] fp32 + split
__global__ void spin_relu_maxpool_kernel(const float* __restrict__ y, int B, float* __restrict__ out, SplitOut outs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 4 channels
    const size_t total = (size_t)B * 56 * 56 * 16;
    if (idx >= total) return;
    const int c4 = (int)(idx % 16);
    const size_t m = idx / 16;
    const int wo = (int)(m % 56), ho = (int)((m / 56) % 56), b = (int)(m / (56 * 56));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);                            // max with 0 = the ReLU
    for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * 2 + kh - 1;
        if (hi < 0 || hi >= 112) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int wi = wo * 2 + kw - 1;
            if (wi < 0 || wi >= 112) continue;
            const float4 v = ld4(y + (((size_t)b * 112 + hi) * 112 + wi) * 64 + c4 * 4);
            r.x = fmaxf(r.x, v.x); r.y = fmaxf(r.y, v.y); r.z = fmaxf(r.z, v.z); r.w = fmaxf(r.w, v.w);
        }
    }
    st4(out + m * 64 + c4 * 4, r);
    store_split4(outs, m * 64 + c4 * 4, r);
}

// im2col of a 3x3 / pad 1 / stride s convolution over a split-bf16 NHWC tensor [B,H,W,C] (C % 8 == 0):
//   A[(b,ho,wo), (kh*3+kw)*C + c] = in[b, ho*s+kh-1, wo*s+kw-1, c]   (zero outside)
__global__ void spin_im2col3_kernel(SplitOut in, int B, int H, int W, int C, int s, int Ho, int Wo, SplitOut A) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 channels of one tap
    const int c8n = C / 8;
    const size_t total = (size_t)B * Ho * Wo * 9 * c8n;
    if (idx >= total) return;
    const int c8 = (int)(idx % c8n);
    const int tap = (int)((idx / c8n) % 9);
    const size_t m = idx / ((size_t)9 * c8n);
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), b = (int)(m / ((size_t)Wo * Ho));
    const int hi = ho * s + tap / 3 - 1, wi = wo * s + tap % 3 - 1;
    uint4 hh = make_uint4(0, 0, 0, 0), ll = hh;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        const size_t src = (((size_t)b * H + hi) * W + wi) * C + c8 * 8;
        hh = *reinterpret_cast<const uint4*>(in.hi + src);
        ll = *reinterpret_cast<const uint4*>(in.lo + src);
    }
    const size_t dst = m * (size_t)(9 * C) + (size_t)tap * C + c8 * 8;
    *reinterpret_cast<uint4*>(A.hi + dst) = hh;
    *reinterpret_cast<uint4*>(A.lo + dst) = ll;
}

// rows (b, ho*s, wo*s) of a split-bf16 NHWC tensor (the A operand of a strided 1x1 downsample convolution, spin.py:116-120)
__global__ void spin_subsample_kernel(SplitOut in, int B, int H, int W, int C, int s, int Ho, int Wo, SplitOut out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c8n = C / 8;
    const size_t total = (size_t)B * Ho * Wo * c8n;
    if (idx >= total) return;
    const int c8 = (int)(idx % c8n);
    const size_t m = idx / c8n;
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), b = (int)(m / ((size_t)Wo * Ho));
    const size_t src = (((size_t)b * H + ho * s) * W + wo * s) * C + c8 * 8;
    *reinterpret_cast<uint4*>(out.hi + m * C + c8 * 8) = *reinterpret_cast<const uint4*>(in.hi + src);
    *reinterpret_cast<uint4*>(out.lo + m * C + c8 * 8) = *reinterpret_cast<const uint4*>(in.lo + src);
}

// y <- relu(y) in place (the next block's residual) + split-bf16 copy (the next block's A operand)   (spin.py:54-55)
__global__ void spin_relu_split_kernel(float* __restrict__ y, size_t n4, SplitOut ys) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n4) return;
    float4 v = ld4(y + idx * 4);
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    st4(y + idx * 4, v);
    store_split4(ys, idx * 4, v);
}

// AvgPool2d(7) over the 7x7 map (spin.py:141-142): x [B,49,C] fp32 (already ReLU'd) -> [B,C]
__global__ void spin_avgpool_kernel(const float* __restrict__ x, int B, int C, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * C) return;
    const int c = idx % C, b = idx / C;
    float acc = 0.f;
    for (int p = 0; p < 49; ++p) acc += x[((size_t)b * 49 + p) * C + c];
    out[idx] = acc * (1.0f / 49.0f);
}
