// Tensor-core attention for SHORT sequences packed into 128-row tiles (the lifter's spatial / temporal self-attention,
// head_dim 64): one tile = G = floor(128 / L) whole sequences of L tokens, one head.
//   S = (Q*scale) K^T  : tcgen05 128 x 128 x 64, split-bf16 (3 MMAs per k16 step), accumulator in TMEM
//   P = exp(S - rowmax) restricted to the row's own sequence (block-diagonal mask), unnormalised, split-bf16 -> smem
//   O = P V            : tcgen05 128 x 64 x 128; V is consumed in place as an MN-major B operand (no transpose);
//                        rows are scaled by 1/rowsum in the epilogue.
// Warp-specialised and persistent over (tile, head) work items. The per-item chain MMA -> softmax -> MMA -> epilogue is
// serial, so TWO independent math groups (4 warps each, own smem buffer, own TMEM accumulators, own barriers) work on
// alternate items and interleave on the SM's four schedulers:
//   warps 0-3 / 4-7 (math group 0 / 1): thread r owns tile row r (= TMEM lane r) for softmax and the epilogue; the
//       group's first thread issues its MMAs. P overwrites the group's Q/K tiles once S is complete.
//   warps 8-15 (loaders): thread (r, half) loads half of q/k/v row r of the next item straight from the fp32 qkv
//       activations (any row addressing: the temporal pass strides over frames), splits to bf16 hi/lo and writes the
//       128-byte-swizzled K-major tiles UMMA expects; arrives on full[group].
// The block-diagonal trick spends 128/L x more MMA flops than the attention needs, which is free here: the CUDA-core
// kernel it replaces ran with 17 of 32 lanes active and ~3 warps per scheduler.
#pragma once
#include <stdlib.h>
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"

constexpr int AT_TILE = 128 * 128;                 // bytes: [128 rows][64 bf16]
constexpr int AT_BUF = 6 * AT_TILE;                // Qh Ql Kh Kl Vh Vl   (P hi aliases Qh|Ql, P lo aliases Kh|Kl)
constexpr int AT_SMEM = 2 * AT_BUF + 1024 + 128;
constexpr int AT_THREADS = 512;

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
__device__ __forceinline__ void bar_sync_group(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
}  // namespace tc

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tile_tc_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, AttnAddr a, SplitOut Os, AttnAddr ao,
                    int L, int G, int nseq, int H, float scale, int dbg) {
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * AT_BUF);   // [2] per group, 256 loader arrivals each
    uint64_t* empty_bar = full_bar + 2;                                    // [2] per group, 1 arrival (tcgen05.commit)
    uint64_t* bar_s = empty_bar + 2;                                       // [2] S complete
    uint64_t* bar_o = bar_s + 2;                                           // [2] O complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_o + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&full_bar[b], 256); tc::mbar_init(&empty_bar[b], 1);
            tc::mbar_init(&bar_s[b], 1); tc::mbar_init(&bar_o[b], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int ntiles = (nseq + G - 1) / G;
    const int nwork = ntiles * H;
    // this CTA's items: work = blockIdx.x + n * gridDim.x, n = 0, 1, ...; item n belongs to group n & 1

    if (warp >= 8) {
        // ================= loaders: 8 warps, thread = (row, half of the 64-wide head slice) =================
        const int lt = tid - 256;
        const int lr = lt & 127, hf = lt >> 7;
        const int lg = lr / L, ltok = lr - lg * L;
        uint32_t n = 0;
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++n) {
            const int grp = n & 1;
            const uint32_t j = n >> 1;                                  // per-group item counter
            const uint32_t base_s = sb + grp * AT_BUF;
            const int tile = work / H, h = work - tile * H;
            const int s = tile * G + lg;
            const bool valid = (lg < G) && (s < nseq) && !(dbg & 1);
            const size_t base = valid ? (size_t)(a.seq(s) + (long long)ltok * a.tok) * a.ld + h * D + hf * 32 : 0;
            float4 x0[8];
            // q first (in flight while we wait for the buffer), then k, v
            if (valid) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x0[i] = ld4(Q + base + i * 4);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tc::mbar_wait(&empty_bar[grp], (j & 1) ^ 1);
#pragma unroll 1
            for (int which = 0; which < 3; ++which) {
                float4 x1[8];
                if (which < 2) {                                         // prefetch the next tensor's row half
                    const float* nxt = which == 0 ? K : V;
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) x1[i] = ld4(nxt + base + i * 4);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) x1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                const float sc = which == 0 ? scale : 1.0f;
                const uint32_t th = base_s + (2 * which) * AT_TILE, tl = th + AT_TILE;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
                    uint4 hh, ll;
                    x[0] = x0[2 * c].x * sc; x[1] = x0[2 * c].y * sc; x[2] = x0[2 * c].z * sc; x[3] = x0[2 * c].w * sc;
                    x[4] = x0[2 * c + 1].x * sc; x[5] = x0[2 * c + 1].y * sc; x[6] = x0[2 * c + 1].z * sc; x[7] = x0[2 * c + 1].w * sc;
                    tc::split8(x, hh, ll);
                    tc::sts16(th, lr, hf * 4 + c, hh); tc::sts16(tl, lr, hf * 4 + c, ll);
                }
                if (which < 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x0[i] = x1[i];
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(&full_bar[grp]);
        }
    } else {
        // ================= math: two groups of 4 warps =================
        const int grp = warp >> 2;
        const int r = tid & 127;                       // tile row owned by this thread
        const int wq = warp & 3;                       // TMEM lane group
        const int g = r / L, tok = r - g * L;
        const uint32_t tS = tmem_base + grp * 256, tO = tS + 128;
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        const uint32_t base_s = sb + grp * AT_BUF;
        const uint32_t Qh = base_s, Ql = base_s + AT_TILE, Kh = base_s + 2 * AT_TILE, Kl = base_s + 3 * AT_TILE, Vh = base_s + 4 * AT_TILE,
                       Vl = base_s + 5 * AT_TILE;
        const uint32_t Ph = Qh, Pl = Kh;               // two [128][64] tiles each, reused once S is complete
        const bool issuer = (tid & 127) == 0;
        // key window of this warp's 32 rows (whole sequences), in 32-column chunks
        const int g_lo = (wq * 32) / L;
        int g_hi = (wq * 32 + 31) / L;
        if (g_hi > G - 1) g_hi = G - 1;
        const int win_lo = g_lo * L, win_hi = (g_hi + 1) * L;           // may be empty for padding-only warps
        uint32_t j = 0;
        for (int work = blockIdx.x + grp * gridDim.x; work < nwork; work += 2 * gridDim.x, ++j) {
            const int tile = work / H, h = work - tile * H;
            const int s = tile * G + g;
            const bool valid = (g < G) && (s < nseq);

            tc::mbar_wait(&full_bar[grp], j & 1);
            if (dbg == 2) {                                                 // profiling knob: loader-only timing
                tc::bar_sync_group(1 + grp);
                if (issuer) tc::mbar_arrive(&empty_bar[grp]);
                continue;
            }
            tc::tc_fence_before();
            tc::bar_sync_group(1 + grp);                                   // the group has finished reading TMEM of its previous item
            if (issuer) {
                tc::tc_fence_after();
                constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 128);
                const uint64_t qh = tc::umma_desc_sw128(Qh), ql = tc::umma_desc_sw128(Ql), kh = tc::umma_desc_sw128(Kh), kl = tc::umma_desc_sw128(Kl);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_bf16(tS, tc::umma_desc_advance_k(ql, k), tc::umma_desc_advance_k(kh, k), idesc, k != 0);
                    tc::umma_bf16(tS, tc::umma_desc_advance_k(qh, k), tc::umma_desc_advance_k(kl, k), idesc, 1);
                    tc::umma_bf16(tS, tc::umma_desc_advance_k(qh, k), tc::umma_desc_advance_k(kh, k), idesc, 1);
                }
                tc::umma_commit(&bar_s[grp]);
            }
            tc::mbar_wait(&bar_s[grp], j & 1);
            tc::tc_fence_after();

            // ---- softmax over this row's own sequence: columns [cs, ce) ----
            const int cs = valid ? g * L : 0, ce = valid ? g * L + L : 0;
            float m = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                if (dbg & 4) break;
                if (c * 32 >= win_hi || c * 32 + 32 <= win_lo) continue;      // warp-uniform
                uint32_t v[32];
                tc::tmem_ld_32x32(tS + lane_sel + c * 32, v);
                tc::tmem_ld_wait();
                if (c * 32 < ce && c * 32 + 32 > cs) {                        // this row's sequence touches the chunk (divergent, cheap)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = c * 32 + i;
                        if (col >= cs && col < ce) m = fmaxf(m, __uint_as_float(v[i]));
                    }
                }
                __syncwarp();
            }
            float lsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                if (dbg & 4) { lsum = 1.f; break; }
                const int t = c >> 1;                                         // P tile (keys 0-63 / 64-127)
                if (c * 32 >= win_hi || c * 32 + 32 <= win_lo) {              // warp-uniform: nothing of this warp's rows lives here
                    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) { tc::sts16(Ph + t * AT_TILE, r, (c & 1) * 4 + jj, z); tc::sts16(Pl + t * AT_TILE, r, (c & 1) * 4 + jj, z); }
                    continue;
                }
                uint32_t v[32];
                tc::tmem_ld_32x32(tS + lane_sel + c * 32, v);
                tc::tmem_ld_wait();
                if (c * 32 < ce && c * 32 + 32 > cs) {                        // this row's sequence touches the chunk
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = c * 32 + jj * 8 + i;
                            // exp2-based fast exponential: |rel err| ~ 2^-21, two orders below the bf16x3 product error
                            p[i] = (col >= cs && col < ce) ? __expf(__uint_as_float(v[jj * 8 + i]) - m) : 0.f;
                            lsum += p[i];
                        }
                        uint4 hh, ll;
                        tc::split8(p, hh, ll);
                        tc::sts16(Ph + t * AT_TILE, r, (c & 1) * 4 + jj, hh);
                        tc::sts16(Pl + t * AT_TILE, r, (c & 1) * 4 + jj, ll);
                    }
                } else {
                    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) { tc::sts16(Ph + t * AT_TILE, r, (c & 1) * 4 + jj, z); tc::sts16(Pl + t * AT_TILE, r, (c & 1) * 4 + jj, z); }
                }
                __syncwarp();
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::bar_sync_group(1 + grp);
            if (issuer) {
                tc::tc_fence_after();
                constexpr uint32_t idesc = tc::umma_idesc_bf16_f32_bmn(128, D);
                const uint64_t vh = tc::umma_desc_sw128(Vh), vl = tc::umma_desc_sw128(Vl);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t ph = tc::umma_desc_advance_k(tc::umma_desc_sw128(Ph + (k >> 2) * AT_TILE), k & 3);
                    const uint64_t pl = tc::umma_desc_advance_k(tc::umma_desc_sw128(Pl + (k >> 2) * AT_TILE), k & 3);
                    const uint64_t bvh = vh + (uint64_t)(k * 2048 >> 4), bvl = vl + (uint64_t)(k * 2048 >> 4);   // 16 keys = two 8-row groups
                    tc::umma_bf16(tO, pl, bvh, idesc, k != 0);
                    tc::umma_bf16(tO, ph, bvl, idesc, 1);
                    tc::umma_bf16(tO, ph, bvh, idesc, 1);
                }
                tc::umma_commit(&bar_o[grp]);
                tc::umma_commit(&empty_bar[grp]);                           // smem buffer is free once these MMAs have read it
            }
            tc::mbar_wait(&bar_o[grp], j & 1);
            tc::tc_fence_after();

            // ---- epilogue: O row / rowsum -> split bf16 -> global ----
            const float inv = 1.0f / lsum;
            const size_t ob = valid ? (size_t)(ao.seq(s) + (long long)tok * ao.tok) * ao.ld + h * D : 0;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v0[32];
                tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v0);
                tc::tmem_ld_wait();
                if (valid && !(dbg & 8)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        store_split4(Os, ob + hf * 32 + i, make_float4(__uint_as_float(v0[i]) * inv, __uint_as_float(v0[i + 1]) * inv,
                                                                       __uint_as_float(v0[i + 2]) * inv, __uint_as_float(v0[i + 3]) * inv));
                }
                __syncwarp();
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
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
    const long long want = (work + 1) / 2;                 // two items in flight per CTA
    const int grid = (int)(want < sms ? (want < 1 ? 1 : want) : sms);
    static int dbg = -1;   // PMCE_ATTN_DEBUG=1: loaders skip global loads; =2: math groups skip all work (profiling only, wrong results)
    if (dbg < 0) { const char* e = getenv("PMCE_ATTN_DEBUG"); dbg = e ? atoi(e) : 0; }
    attn_tile_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(Q, K, V, a, Os, ao, L, G, nseq, H, 1.0f / sqrtf(64.0f), dbg);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
