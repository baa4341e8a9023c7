// Tensor-core attention for SHORT sequences packed into 128-row tiles (the lifter's spatial / temporal self-attention,
// head_dim 64): one tile = G = floor(128 / L) whole sequences of L tokens, one head.
//   S = (Q*scale) K^T  : tcgen05 128 x 128 x 64, split-bf16 (3 MMAs per k16 step), accumulator in TMEM columns [0,128)
//   P = exp(S - rowmax) restricted to the row's own sequence (block-diagonal mask), unnormalised, split-bf16 -> smem
//   O = P V            : tcgen05 128 x 64 x 128; V is consumed in place as an MN-major B operand (no transpose),
//                        accumulator in TMEM columns [128,192); rows are scaled by 1/rowsum in the epilogue.
// 128 threads: thread r owns tile row r (= TMEM lane r) for loading, softmax and the epilogue; thread 0 issues the MMAs.
// Q/K/V come straight from the fp32 qkv activations (any row addressing: the temporal pass strides over frames), are
// split to bf16 hi/lo on the fly and written to shared memory in the 128-byte-swizzled K-major layout UMMA expects.
// The block-diagonal trick spends 128/L x more MMA flops than the attention needs, which is free here: the CUDA-core
// kernel it replaces ran at ~15% lane/issue efficiency with 17 of 32 lanes active.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"

constexpr int AT_TILE = 128 * 128;                 // bytes: [128 rows][64 bf16]
constexpr int AT_NTILES = 10;                      // Qh Ql Kh Kl Vh Vl Ph0 Ph1 Pl0 Pl1
constexpr int AT_SMEM = AT_NTILES * AT_TILE + 1024 + 64;

namespace tc {
__device__ __forceinline__ void sts16(uint32_t tile, int row, int chunk, uint4 v) {   // 16-byte chunk, SW128 pattern
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile + row * 128 + ((chunk ^ (row & 7)) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void split8(const float (&x)[8], uint4& h, uint4& l) {
    split_bf16x2(x[0], x[1], h.x, l.x); split_bf16x2(x[2], x[3], h.y, l.y);
    split_bf16x2(x[4], x[5], h.z, l.z); split_bf16x2(x[6], x[7], h.w, l.w);
}
// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_bmn(int M, int N) { return umma_idesc_bf16_f32(M, N) | (1u << 16); }
}  // namespace tc

__global__ void __launch_bounds__(128, 1)
attn_tile_tc_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, AttnAddr a, SplitOut Os, AttnAddr ao,
                    int L, int G, int nseq, int H, float scale) {
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t Qh = sb, Ql = sb + AT_TILE, Kh = sb + 2 * AT_TILE, Kl = sb + 3 * AT_TILE, Vh = sb + 4 * AT_TILE, Vl = sb + 5 * AT_TILE;
    const uint32_t Ph = sb + 6 * AT_TILE, Pl = sb + 8 * AT_TILE;          // two [128][64] tiles each (keys 0-63, 64-127)
    uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + AT_NTILES * AT_TILE);
    uint64_t* bar_o = bar_s + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_o + 1);

    const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
    if (r == 0) {
        tc::mbar_init(bar_s, 1);
        tc::mbar_init(bar_o, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 256);
    // P tiles start as zeros; each row only ever rewrites the key window of its warp (same window every iteration)
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) { tc::sts16(Ph + t * AT_TILE, r, c, z); tc::sts16(Pl + t * AT_TILE, r, c, z); }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

    // key window of this warp's 32 rows (whole sequences), in 32-column chunks
    const int g_lo = (warp * 32) / L;
    int g_hi = (warp * 32 + 31) / L;
    if (g_hi > G - 1) g_hi = G - 1;
    const int win_lo = g_lo * L, win_hi = (g_hi + 1) * L;               // may be empty (win_lo >= win_hi) for padding-only warps
    const int g = r / L, tok = r - g * L;
    const int ntiles = (nseq + G - 1) / G;
    const int nwork = ntiles * H;
    uint32_t iter = 0;

    for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++iter) {
        const int tile = work / H, h = work - tile * H;
        const int s = tile * G + g;
        const bool valid = (g < G) && (s < nseq);
        // ---- load q, k, v rows (fp32) -> split bf16 -> swizzled K-major tiles ----
        if (valid) {
            const size_t base = (size_t)(a.seq(s) + (long long)tok * a.tok) * a.ld + h * D;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float x[8];
                uint4 hh, ll;
                float4 u = ld4(Q + base + c * 8), w = ld4(Q + base + c * 8 + 4);
                x[0] = u.x * scale; x[1] = u.y * scale; x[2] = u.z * scale; x[3] = u.w * scale;
                x[4] = w.x * scale; x[5] = w.y * scale; x[6] = w.z * scale; x[7] = w.w * scale;
                tc::split8(x, hh, ll);
                tc::sts16(Qh, r, c, hh); tc::sts16(Ql, r, c, ll);
                u = ld4(K + base + c * 8); w = ld4(K + base + c * 8 + 4);
                x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = w.x; x[5] = w.y; x[6] = w.z; x[7] = w.w;
                tc::split8(x, hh, ll);
                tc::sts16(Kh, r, c, hh); tc::sts16(Kl, r, c, ll);
                u = ld4(V + base + c * 8); w = ld4(V + base + c * 8 + 4);
                x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = w.x; x[5] = w.y; x[6] = w.z; x[7] = w.w;
                tc::split8(x, hh, ll);
                tc::sts16(Vh, r, c, hh); tc::sts16(Vl, r, c, ll);
            }
        } else {
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                tc::sts16(Qh, r, c, z); tc::sts16(Ql, r, c, z); tc::sts16(Kh, r, c, z);
                tc::sts16(Kl, r, c, z); tc::sts16(Vh, r, c, z); tc::sts16(Vl, r, c, z);
            }
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (r == 0) {
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 128);
            const uint64_t qh = tc::umma_desc_sw128(Qh), ql = tc::umma_desc_sw128(Ql), kh = tc::umma_desc_sw128(Kh), kl = tc::umma_desc_sw128(Kl);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(tS, tc::umma_desc_advance_k(ql, k), tc::umma_desc_advance_k(kh, k), idesc, k != 0);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(qh, k), tc::umma_desc_advance_k(kl, k), idesc, 1);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(qh, k), tc::umma_desc_advance_k(kh, k), idesc, 1);
            }
            tc::umma_commit(bar_s);
        }
        tc::mbar_wait(bar_s, iter & 1);
        tc::tc_fence_after();

        // ---- softmax over this row's own sequence: columns [cs, ce) ----
        const int cs = valid ? g * L : 0, ce = valid ? g * L + L : 0;
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            if (c * 32 >= win_hi || c * 32 + 32 <= win_lo) continue;      // warp-uniform
            uint32_t v[32];
            tc::tmem_ld_32x32(tS + lane_sel + c * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int col = c * 32 + i;
                if (col >= cs && col < ce) m = fmaxf(m, __uint_as_float(v[i]));
            }
        }
        float lsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            if (c * 32 >= win_hi || c * 32 + 32 <= win_lo) continue;      // warp-uniform
            uint32_t v[32];
            tc::tmem_ld_32x32(tS + lane_sel + c * 32, v);
            tc::tmem_ld_wait();
            const int t = c >> 1;                                         // P tile (keys 0-63 / 64-127)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = c * 32 + j * 8 + i;
                    p[i] = (col >= cs && col < ce) ? expf(__uint_as_float(v[j * 8 + i]) - m) : 0.f;
                    lsum += p[i];
                }
                uint4 hh, ll;
                tc::split8(p, hh, ll);
                const int chunk = (c & 1) * 4 + j;
                tc::sts16(Ph + t * AT_TILE, r, chunk, hh);
                tc::sts16(Pl + t * AT_TILE, r, chunk, ll);
            }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (r == 0) {
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32_bmn(128, D);
            const uint64_t vh = tc::umma_desc_sw128(Vh), vl = tc::umma_desc_sw128(Vl);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint64_t ph = tc::umma_desc_advance_k(tc::umma_desc_sw128(Ph + (j >> 2) * AT_TILE), j & 3);
                const uint64_t pl = tc::umma_desc_advance_k(tc::umma_desc_sw128(Pl + (j >> 2) * AT_TILE), j & 3);
                const uint64_t bvh = vh + (uint64_t)(j * 2048 >> 4), bvl = vl + (uint64_t)(j * 2048 >> 4);   // 16 keys = two 8-row groups
                tc::umma_bf16(tO, pl, bvh, idesc, j != 0);
                tc::umma_bf16(tO, ph, bvl, idesc, 1);
                tc::umma_bf16(tO, ph, bvh, idesc, 1);
            }
            tc::umma_commit(bar_o);
        }
        tc::mbar_wait(bar_o, iter & 1);
        tc::tc_fence_after();

        // ---- epilogue: O row / rowsum -> split bf16 -> global ----
        {
            uint32_t v0[32], v1[32];
            tc::tmem_ld_32x32(tO + lane_sel, v0);
            tc::tmem_ld_32x32(tO + lane_sel + 32, v1);
            tc::tmem_ld_wait();
            if (valid) {
                const float inv = 1.0f / lsum;
                const size_t ob = (size_t)(ao.seq(s) + (long long)tok * ao.tok) * ao.ld + h * D;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    store_split4(Os, ob + i, make_float4(__uint_as_float(v0[i]) * inv, __uint_as_float(v0[i + 1]) * inv,
                                                         __uint_as_float(v0[i + 2]) * inv, __uint_as_float(v0[i + 3]) * inv));
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    store_split4(Os, ob + 32 + i, make_float4(__uint_as_float(v1[i]) * inv, __uint_as_float(v1[i + 1]) * inv,
                                                              __uint_as_float(v1[i + 2]) * inv, __uint_as_float(v1[i + 3]) * inv));
            }
        }
        tc::tc_fence_before();
        __syncthreads();          // TMEM S/O and the smem tiles are free for the next (tile, head)
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// nseq sequences of L tokens (L <= 128), H heads of 64; fp32 q/k/v addressed by `a`, split-bf16 output by `ao`.
static inline int launch_attn_tile_tc(const float* Q, const float* K, const float* V, AttnAddr a, SplitOut Os, AttnAddr ao, int nseq, int H, int L,
                                      cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(attn_tile_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) return 2;
        configured = true;
    }
    const int G = 128 / L;
    const int ntiles = (nseq + G - 1) / G;
    const long long work = (long long)ntiles * H;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(work < sms ? work : sms);
    attn_tile_tc_kernel<<<grid, 128, AT_SMEM, st>>>(Q, K, V, a, Os, ao, L, G, nseq, H, 1.0f / sqrtf(64.0f));
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
