// Fused vertex<-joint CrossAttentionBlock front (CoevoDecoder.py:47-62 inside :82-86), one kernel per block:
//     xq' = xq + Wp . MHA( Wq AdaLN_q(xq) + bq ,  K ,  V ) + bp            (K, V: the 17/19 projected joint rows of the clip)
//     t   = AdaLN_2(xq')  (split-bf16: the A operand of the block's fc1 GEMM)
// replacing AdaLN -> Wq GEMM -> attention -> Wp GEMM(+residual) -> AdaLN (five launches, each a round trip of the
// [B*431, 64] stream through HBM) with one pass: the stream is read once and written once.
//
// Work item = 128 query rows of one clip (TMA box [1][128][32 fp32] x2 out of the 3-D view [B][431][64]: rows past 431 are
// zero-filled on load and clipped on store). 128 threads, thread r owns tile row r = TMEM lane r:
//   TMA load (SW128)            -> row in registers -> AdaLN_q -> split-bf16 -> swizzled A tiles (hi, lo)
//   tcgen05 128x64x64 (bf16x3)  -> Q in TMEM -> registers (+bq, *scale)
//   attention on CUDA cores     : keys/values of the clip (<= 24 rows) broadcast from shared memory, softmax in registers
//   O -> split-bf16 -> A tiles  -> tcgen05 128x64x64 with Wp -> TMEM -> + bp + xq (re-read from the swizzled input tile)
//   xq' -> input tile (same swizzled slot) -> TMA store; AdaLN_2(xq') -> A tiles -> TMA store (hi, lo)
// The four weight tiles ([64][64] bf16 hi/lo of Wq and Wp) are loaded once per CTA by TMA; CTAs are persistent over items,
// two per SM (110 KB of shared memory, 128 TMEM columns each) so one CTA's loads/stores overlap the other's math.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"

constexpr int CAF_THREADS = 128;
constexpr int CAF_KV_ROWS = 24;                         // >= num_joint (17 h36m, 19 coco)
constexpr int CAF_IN = 2 * 128 * 128;                   // two [128][32 fp32] boxes
constexpr int CAF_OFF_A = CAF_IN;                       // A hi | A lo  ([128][64 bf16] each)
constexpr int CAF_OFF_W = CAF_OFF_A + 2 * AT_TILE;      // Wq hi | Wq lo | Wp hi | Wp lo  ([64][64 bf16] each)
constexpr int CAF_OFF_KV = CAF_OFF_W + 4 * 64 * 128;    // K [24][64] fp32 | V [24][64] fp32
constexpr int CAF_OFF_GB = CAF_OFF_KV + 2 * CAF_KV_ROWS * 64 * 4;   // gamma_q beta_q gamma_2 beta_2 bq bp  (6 x 64 fp32)
constexpr int CAF_OFF_BAR = CAF_OFF_GB + 6 * 64 * 4;
constexpr int CAF_SMEM = CAF_OFF_BAR + 64 + 1024;       // + alignment slack

namespace tc {
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int crd0, int crd1, int crd2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(crd0), "r"(crd1), "r"(crd2)
                 : "memory");
}
__device__ __forceinline__ float4 lds16(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts16f(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
}  // namespace tc

struct CaFusedArgs {
    const float* K;          // [B*N2, 64] projected keys (fp32)
    const float* V;          // [B*N2, 64] projected values
    const float* gb;         // [B, gb_ld] AdaLN gamma/beta of every slot (pmce_adaln_gammabeta)
    const float* bq;         // [64]
    const float* bp;         // [64]
    int gb_ld, slot_q, slot_2;
    int B, N1, N2, qtiles;
    float scale, eps;
};

// AdaLayerNorm of one 64-wide row held in registers (CoevoDecoder.py:23-29): unbiased std, eps added to the std.
__device__ __forceinline__ void caf_adaln_split_store(const float (&x)[64], const float* __restrict__ gam, const float* __restrict__ bet, float eps,
                                                      uint32_t a_hi, uint32_t a_lo, int r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += x[i];
    const float mean = s * (1.0f / 64.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float inv = 1.0f / (sqrtf(q * (1.0f / 63.0f)) + eps);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        const float4 g0 = ld4(gam + cc * 8), g1 = ld4(gam + cc * 8 + 4), b0 = ld4(bet + cc * 8), b1 = ld4(bet + cc * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = gg[i] * (x[cc * 8 + i] - mean) * inv + bb[i];
        uint4 hh, ll;
        tc::split8(y, hh, ll);
        tc::sts16(a_hi, r, cc, hh);
        tc::sts16(a_lo, r, cc, ll);
    }
}

template <int H, int NK>
__global__ void __launch_bounds__(CAF_THREADS, 2)
ca_vertex_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_thi, const __grid_constant__ CUtensorMap tm_tlo,
                       const __grid_constant__ CUtensorMap tm_wq_hi, const __grid_constant__ CUtensorMap tm_wq_lo,
                       const __grid_constant__ CUtensorMap tm_wp_hi, const __grid_constant__ CUtensorMap tm_wp_lo, CaFusedArgs a) {
    constexpr int D = 64 / H;
    static_assert(NK <= CAF_KV_ROWS, "too many keys");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t s_in = sb, s_ahi = sb + CAF_OFF_A, s_alo = s_ahi + AT_TILE, s_w = sb + CAF_OFF_W;
    float* Ks = reinterpret_cast<float*>(smem + CAF_OFF_KV);
    float* Vs = Ks + CAF_KV_ROWS * 64;
    float* gbs = reinterpret_cast<float*>(smem + CAF_OFF_GB);          // [0]=gamma_q [1]=beta_q [2]=gamma_2 [3]=beta_2 [4]=bq [5]=bp
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(smem + CAF_OFF_BAR);
    uint64_t* bar_w = bar_in + 1;
    uint64_t* bar_mma = bar_in + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_in + 3);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tm_x); tc::tma_prefetch_desc(&tm_thi); tc::tma_prefetch_desc(&tm_tlo);
        tc::tma_prefetch_desc(&tm_wq_hi); tc::tma_prefetch_desc(&tm_wq_lo); tc::tma_prefetch_desc(&tm_wp_hi); tc::tma_prefetch_desc(&tm_wp_lo);
        tc::mbar_init(bar_in, 1); tc::mbar_init(bar_w, 1); tc::mbar_init(bar_mma, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 128);
    if (tid < 64) { gbs[4 * 64 + tid] = a.bq[tid]; gbs[5 * 64 + tid] = a.bp[tid]; }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tQ = tmem_base, tP = tmem_base + 64;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

    if (tid == 0) {                                   // the four weight tiles, once per CTA
        tc::mbar_arrive_expect_tx(bar_w, 4 * 64 * 128);
        tc::tma_load_2d(smem + CAF_OFF_W, &tm_wq_hi, bar_w, 0, 0);
        tc::tma_load_2d(smem + CAF_OFF_W + 8192, &tm_wq_lo, bar_w, 0, 0);
        tc::tma_load_2d(smem + CAF_OFF_W + 16384, &tm_wp_hi, bar_w, 0, 0);
        tc::tma_load_2d(smem + CAF_OFF_W + 24576, &tm_wp_lo, bar_w, 0, 0);
    }

    const int r = tid;
    const uint32_t row_in0 = s_in + r * 128, row_in1 = s_in + 16384 + r * 128;
    const int sw = r & 7;
    const int ntiles = a.B * a.qtiles;
    uint32_t it = 0, mph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;
        if (tid == 0) {
            if (it > 0) tc::tma_store_wait_read<0>();          // the previous item's stores have read the IN / A tiles
            tc::mbar_arrive_expect_tx(bar_in, CAF_IN);
            tc::tma_load_3d(smem, &tm_x, bar_in, 0, row0, b);
            tc::tma_load_3d(smem + 16384, &tm_x, bar_in, 32, row0, b);
        }
        // keys / values / AdaLN parameters of clip b
        for (int idx = tid; idx < NK * 16; idx += CAF_THREADS) {
            const int kr = idx >> 4, c = (idx & 15) * 4;
            st4(Ks + kr * 64 + c, ld4(a.K + ((size_t)b * a.N2 + kr) * 64 + c));
            st4(Vs + kr * 64 + c, ld4(a.V + ((size_t)b * a.N2 + kr) * 64 + c));
        }
        {
            const float* g = a.gb + (size_t)b * a.gb_ld;
            const int which = tid >> 6, c = tid & 63;              // threads 0-63: slot_q, 64-127: slot_2 (gamma | beta)
            const float* src = g + (which == 0 ? a.slot_q : a.slot_2) * 128;
            gbs[(which * 2) * 64 + c] = src[c];
            gbs[(which * 2 + 1) * 64 + c] = src[64 + c];
        }
        __syncthreads();

        // ---- xq row -> AdaLN_q -> split A tiles ----
        tc::mbar_wait(bar_in, it & 1);
        {
            float x[64];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v0 = tc::lds16(row_in0 + ((j ^ sw) << 4)), v1 = tc::lds16(row_in1 + ((j ^ sw) << 4));
                x[4 * j] = v0.x; x[4 * j + 1] = v0.y; x[4 * j + 2] = v0.z; x[4 * j + 3] = v0.w;
                x[32 + 4 * j] = v1.x; x[32 + 4 * j + 1] = v1.y; x[32 + 4 * j + 2] = v1.z; x[32 + 4 * j + 3] = v1.w;
            }
            caf_adaln_split_store(x, gbs, gbs + 64, a.eps, s_ahi, s_alo, r);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 64);
        if (tid == 0) {
            if (it == 0) tc::mbar_wait(bar_w, 0);
            tc::tc_fence_after();
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w), wl = tc::umma_desc_sw128(s_w + 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(tQ, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                tc::umma_bf16(tQ, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                tc::umma_bf16(tQ, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
            }
            tc::umma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, mph & 1);
        ++mph;
        tc::tc_fence_after();

        // ---- Q row from TMEM; attention over the clip's NK keys on CUDA cores ----
        float o[64];
        {
            float q[64];
            {
                uint32_t v[32];
                tc::tmem_ld_32x32(tQ + lane_sel, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) q[i] = (__uint_as_float(v[i]) + gbs[4 * 64 + i]) * a.scale;
                tc::tmem_ld_32x32(tQ + lane_sel + 32, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) q[32 + i] = (__uint_as_float(v[i]) + gbs[4 * 64 + 32 + i]) * a.scale;
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float s[NK];
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < NK; ++j) {
                    const float* kr = Ks + j * 64 + h * D;
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int c = 0; c < D; c += 4) {
                        const float4 kk = ld4(kr + c);
                        s0 = fmaf(q[h * D + c], kk.x, s0); s1 = fmaf(q[h * D + c + 1], kk.y, s1);
                        s0 = fmaf(q[h * D + c + 2], kk.z, s0); s1 = fmaf(q[h * D + c + 3], kk.w, s1);
                    }
                    s[j] = s0 + s1;
                    m = fmaxf(m, s[j]);
                }
                float l = 0.f;
#pragma unroll
                for (int j = 0; j < NK; ++j) { s[j] = expf(s[j] - m); l += s[j]; }
                const float inv = 1.0f / l;
#pragma unroll
                for (int c = 0; c < D; ++c) o[h * D + c] = 0.f;
#pragma unroll
                for (int j = 0; j < NK; ++j) {
                    const float* vr = Vs + j * 64 + h * D;
                    const float p = s[j] * inv;
#pragma unroll
                    for (int c = 0; c < D; c += 4) {
                        const float4 vv = ld4(vr + c);
                        o[h * D + c] = fmaf(p, vv.x, o[h * D + c]); o[h * D + c + 1] = fmaf(p, vv.y, o[h * D + c + 1]);
                        o[h * D + c + 2] = fmaf(p, vv.z, o[h * D + c + 2]); o[h * D + c + 3] = fmaf(p, vv.w, o[h * D + c + 3]);
                    }
                }
            }
        }
        // ---- O -> split A tiles -> proj MMA ----
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = o[cc * 8 + i];
            uint4 hh, ll;
            tc::split8(y, hh, ll);
            tc::sts16(s_ahi, r, cc, hh);
            tc::sts16(s_alo, r, cc, ll);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w + 16384), wl = tc::umma_desc_sw128(s_w + 24576);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(tP, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                tc::umma_bf16(tP, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                tc::umma_bf16(tP, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
            }
            tc::umma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, mph & 1);
        ++mph;
        tc::tc_fence_after();

        // ---- xq' = (proj + bp) + xq -> input tile; AdaLN_2(xq') -> split A tiles; TMA stores ----
        {
            float x[64];
            {
                uint32_t v[32];
                tc::tmem_ld_32x32(tP + lane_sel, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]) + gbs[5 * 64 + i];
                tc::tmem_ld_32x32(tP + lane_sel + 32, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) x[32 + i] = __uint_as_float(v[i]) + gbs[5 * 64 + 32 + i];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t a0 = row_in0 + ((j ^ sw) << 4), a1 = row_in1 + ((j ^ sw) << 4);
                const float4 v0 = tc::lds16(a0), v1 = tc::lds16(a1);
                x[4 * j] += v0.x; x[4 * j + 1] += v0.y; x[4 * j + 2] += v0.z; x[4 * j + 3] += v0.w;
                x[32 + 4 * j] += v1.x; x[32 + 4 * j + 1] += v1.y; x[32 + 4 * j + 2] += v1.z; x[32 + 4 * j + 3] += v1.w;
                tc::sts16f(a0, make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
                tc::sts16f(a1, make_float4(x[32 + 4 * j], x[32 + 4 * j + 1], x[32 + 4 * j + 2], x[32 + 4 * j + 3]));
            }
            caf_adaln_split_store(x, gbs + 128, gbs + 192, a.eps, s_ahi, s_alo, r);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tma_store_3d(&tm_x, smem, 0, row0, b);
            tc::tma_store_3d(&tm_x, smem + 16384, 32, row0, b);
            tc::tma_store_3d(&tm_thi, smem + CAF_OFF_A, 0, row0, b);
            tc::tma_store_3d(&tm_tlo, smem + CAF_OFF_A + AT_TILE, 0, row0, b);
            tc::tma_store_commit();
        }
    }
    if (tid == 0) tc::tma_store_wait<0>();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

// generic 3-D tiled map: dims/box innermost first, element strides ld1 / ld2 of dims 1 / 2, 128-byte swizzle
static inline int make_tmap_3d(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int d0, int d1, int d2, long long ld1,
                               long long ld2, int b0, int b1, int b2) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)ld1 * elem_bytes, (cuuint64_t)ld2 * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

template <int H, int NK>
static inline int launch_ca_vertex_fused_t(const CUtensorMap* maps, const CaFusedArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ca_vertex_fused_kernel<H, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, CAF_SMEM) != cudaSuccess) return 2;
        configured = true;
    }
    const int ntiles = a.B * a.qtiles;
    const int cap = 2 * tc_num_sms();
    const int grid = ntiles < cap ? ntiles : cap;
    ca_vertex_fused_kernel<H, NK><<<grid, CAF_THREADS, CAF_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], a);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// true when the fused kernel covers (heads, keys): the reference's vertex stream has 2 heads; 17 (h36m) / 19 (coco) joints
static inline bool ca_vertex_fused_supported(int heads, int nkeys) { return heads == 2 && (nkeys == 17 || nkeys == 19); }

// xq [B, N1, 64] fp32 (updated in place), t_hi/t_lo [B, N1, 64] bf16 (AdaLN_2 of the result, split);
// wq_* / wp_*: [64, 64] bf16 hi/lo weight copies.
static inline int launch_ca_vertex_fused(float* xq, __nv_bfloat16* t_hi, __nv_bfloat16* t_lo, const __nv_bfloat16* wq_hi, const __nv_bfloat16* wq_lo,
                                         const __nv_bfloat16* wp_hi, const __nv_bfloat16* wp_lo, int heads, CaFusedArgs a, cudaStream_t st) {
    CUtensorMap maps[7];
    a.qtiles = (a.N1 + 127) / 128;
    if (make_tmap_3d(&maps[0], xq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, a.N1, a.B, 64, 64LL * a.N1, 32, 128, 1) ||
        make_tmap_3d(&maps[1], t_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64, a.N1, a.B, 64, 64LL * a.N1, 64, 128, 1) ||
        make_tmap_3d(&maps[2], t_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64, a.N1, a.B, 64, 64LL * a.N1, 64, 128, 1) ||
        make_tmap_bf16(&maps[3], wq_hi, 64, 64, 64, 64) || make_tmap_bf16(&maps[4], wq_lo, 64, 64, 64, 64) ||
        make_tmap_bf16(&maps[5], wp_hi, 64, 64, 64, 64) || make_tmap_bf16(&maps[6], wp_lo, 64, 64, 64, 64))
        return 1;
    if (heads == 2 && a.N2 == 17) return launch_ca_vertex_fused_t<2, 17>(maps, a, st);
    if (heads == 2 && a.N2 == 19) return launch_ca_vertex_fused_t<2, 19>(maps, a, st);
    return 4;
}
