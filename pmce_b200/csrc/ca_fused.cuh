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

constexpr int CAF_THREADS = 256;
constexpr int CAF_KP = 32;                              // key slots per head (num_joint <= 24 live: 17 h36m, 19 coco; the rest are zero)
constexpr int CAF_MAXJ = 24;
constexpr int CAF_H = 2;                                // heads of the vertex stream (CoevoDecoder.py:140)
constexpr int CAF_NS = CAF_H * CAF_KP;                  // 64 score columns: column 32 h + j = (head h, key j)
constexpr int CAF_IN = 2 * 128 * 128;                   // two [128][32 fp32] boxes; the SAME 32 KB then hold the A tiles (hi | lo)
constexpr int CAF_OFF_W = CAF_IN;                       // KQ hi | KQ lo | VPt hi | VPt lo ([64][64 bf16] each); later the t tiles (hi | lo)
constexpr int CAF_OFF_GB = CAF_OFF_W + 4 * 8192;        // gamma_q beta_q gamma_2 beta_2 bp  (5 x 64 fp32)
constexpr int CAF_OFF_PART = CAF_OFF_GB + 5 * 64 * 4;   // per-thread partial sums of the LayerNorm statistics (256 fp32)
constexpr int CAF_OFF_BAR = CAF_OFF_PART + 256 * 4;
constexpr int CAF_SMEM = CAF_OFF_BAR + 64 + 1024;       // + alignment slack  (68 KB: three CTAs per SM)
constexpr int CAF_TX = CAF_IN + 2 * CAF_NS * 128 + 2 * 64 * 128;   // bytes per item arriving on bar_in

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
    const float* sb;         // [B, 64] folded score bias  scale * bq_h . K_hj
    const float* gb;         // [B, gb_ld] AdaLN gamma/beta of every slot (pmce_adaln_gammabeta)
    const float* bp;         // [64]
    int gb_ld, slot_q, slot_2;
    int B, N1, N2, qtiles;
    float eps;
};

// read the 32 floats of one half row out of a swizzled [128][32 fp32] box
__device__ __forceinline__ void caf_read_half(uint32_t rowaddr, int sw, float (&x)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = tc::lds16(rowaddr + ((j ^ sw) << 4));
        x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
}

// AdaLayerNorm (CoevoDecoder.py:23-29: unbiased std, eps added to the std) of a 64-wide row owned by TWO threads (32 columns
// each, in different warps): two-pass statistics, the halves' partial sums meet in shared memory (`part`, one float per
// thread; partner = tid ^ 128). Contains two __syncthreads, the first of which also orders every thread's earlier shared-
// memory reads before the tile writes below. Writes the normalised half as split-bf16 into swizzled [128][64 bf16] tiles.
__device__ __forceinline__ void caf_adaln_pair(const float (&own)[32], float* part, int tid, const float* __restrict__ gam,
                                               const float* __restrict__ bet, float eps, uint32_t t_hi, uint32_t t_lo, int r, int hf) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) { s0 += own[i]; s1 += own[i + 1]; }
    part[tid] = s0 + s1;
    __syncthreads();
    const float mean = ((s0 + s1) + part[tid ^ 128]) * (1.0f / 64.0f);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const float d0 = own[i] - mean, d1 = own[i + 1] - mean;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
    }
    __syncthreads();                                   // every partner has read the sums before `part` is reused
    part[tid] = q0 + q1;
    __syncthreads();
    const float inv = 1.0f / (sqrtf(((q0 + q1) + part[tid ^ 128]) * (1.0f / 63.0f)) + eps);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = hf * 32 + cc * 8;
        const float4 g0 = ld4(gam + c0), g1 = ld4(gam + c0 + 4), b0 = ld4(bet + c0), b1 = ld4(bet + c0 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = gg[i] * (own[cc * 8 + i] - mean) * inv + bb[i];
        uint4 hh, ll;
        tc::split8(y, hh, ll);
        tc::sts16(t_hi, r, hf * 4 + cc, hh);
        tc::sts16(t_lo, r, hf * 4 + cc, ll);
    }
}

// Few-key cross-attention as two skinny GEMMs. With K_h, V_h the clip's projected keys/values of head h (NK <= 24 rows):
//   scores_h = scale (xn Wq_h^T + bq_h) K_h^T = xn (scale K_h Wq_h)^T + scale K_h bq_h         -> KQ [64][64], sb [64]
//   proj(concat_h P_h V_h) = sum_h P_h (V_h Wp[:, h]^T) + bp                                    -> VPt [64][64]
// (row / column 32 h + j = head h, key j; slots j >= NK are zero) so per 128-row item: S = AdaLN_q(xq) KQ^T (tcgen05
// 128x64x64), softmax per head in registers, out = P VPt^T (tcgen05 128x64x64). KQ / VPt / sb are per clip, made by
// ca_joint_fold_kernel (split-bf16), and arrive by TMA.
// Shared memory is two 32 KB regions that change roles through the item: X = {xq fp32 boxes -> A tiles (AdaLN_q(xq) split)
// -> P split -> xq' fp32 boxes for the TMA store}, W = {KQ | VPt -> t tiles (AdaLN_2(xq') split) for the TMA store}; the
// thread keeps its 32 xq values in registers for the residual. 68 KB and <= 80 registers: three CTAs (24 warps) per SM,
// which is what hides the TMA / MMA / TMEM round trips of the per-item chain.
template <int NK>
__global__ void __launch_bounds__(CAF_THREADS, 3)
ca_vertex_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_thi, const __grid_constant__ CUtensorMap tm_tlo,
                       const __grid_constant__ CUtensorMap tm_kq_hi, const __grid_constant__ CUtensorMap tm_kq_lo,
                       const __grid_constant__ CUtensorMap tm_vp_hi, const __grid_constant__ CUtensorMap tm_vp_lo, CaFusedArgs a) {
    static_assert(NK <= CAF_MAXJ, "too many keys");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t s_ahi = sb, s_alo = sb + AT_TILE, s_w = sb + CAF_OFF_W;
    float* gbs = reinterpret_cast<float*>(smem + CAF_OFF_GB);          // [0]=gamma_q [1]=beta_q [2]=gamma_2 [3]=beta_2 [4]=bp
    float* part = reinterpret_cast<float*>(smem + CAF_OFF_PART);
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(smem + CAF_OFF_BAR);
    uint64_t* bar_mma = bar_in + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_in + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tm_x); tc::tma_prefetch_desc(&tm_thi); tc::tma_prefetch_desc(&tm_tlo);
        tc::tma_prefetch_desc(&tm_kq_hi); tc::tma_prefetch_desc(&tm_kq_lo); tc::tma_prefetch_desc(&tm_vp_hi); tc::tma_prefetch_desc(&tm_vp_lo);
        tc::mbar_init(bar_in, 1); tc::mbar_init(bar_mma, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        if ((int)blockIdx.x < a.B * a.qtiles) {        // the first item's loads fly while TMEM is allocated and the CTA assembles
            const int b = blockIdx.x / a.qtiles, row0 = (blockIdx.x % a.qtiles) * 128;
            tc::mbar_arrive_expect_tx(bar_in, CAF_TX);
            tc::tma_load_3d(smem, &tm_x, bar_in, 0, row0, b);
            tc::tma_load_3d(smem + 16384, &tm_x, bar_in, 32, row0, b);
            tc::tma_load_2d(smem + CAF_OFF_W, &tm_kq_hi, bar_in, 0, b * CAF_NS);
            tc::tma_load_2d(smem + CAF_OFF_W + 8192, &tm_kq_lo, bar_in, 0, b * CAF_NS);
            tc::tma_load_2d(smem + CAF_OFF_W + 16384, &tm_vp_hi, bar_in, 0, b * 64);
            tc::tma_load_2d(smem + CAF_OFF_W + 24576, &tm_vp_lo, bar_in, 0, b * 64);
        }
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 128);
    if (tid < 64) gbs[4 * 64 + tid] = a.bp[tid];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tS = tmem_base, tO = tmem_base + 64;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;

    // thread = (tile row r, column half hf): hf selects the 32-column half of the 64-wide row it owns and the attention head
    // whose softmax it runs
    const int r = tid & 127, hf = tid >> 7;
    const uint32_t row_own = sb + hf * 16384 + r * 128;                // this thread's half row inside the fp32 boxes
    const int sw = r & 7;
    const int ntiles = a.B * a.qtiles;
    uint32_t it = 0, mph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;
        if (tid == 0 && it > 0) {                              // (the first item's loads were issued in the prologue)
            tc::tma_store_wait_read<0>();                      // the previous item's stores have read both regions
            tc::mbar_arrive_expect_tx(bar_in, CAF_TX);
            tc::tma_load_3d(smem, &tm_x, bar_in, 0, row0, b);
            tc::tma_load_3d(smem + 16384, &tm_x, bar_in, 32, row0, b);
            tc::tma_load_2d(smem + CAF_OFF_W, &tm_kq_hi, bar_in, 0, b * CAF_NS);
            tc::tma_load_2d(smem + CAF_OFF_W + 8192, &tm_kq_lo, bar_in, 0, b * CAF_NS);
            tc::tma_load_2d(smem + CAF_OFF_W + 16384, &tm_vp_hi, bar_in, 0, b * 64);
            tc::tma_load_2d(smem + CAF_OFF_W + 24576, &tm_vp_lo, bar_in, 0, b * 64);
        }
        {   // AdaLN parameters of clip b: gamma|beta of slot_q and slot_2 (the previous item's readers are past its last barrier)
            const int arr = tid >> 6, c = tid & 63;
            gbs[arr * 64 + c] = a.gb[(size_t)b * a.gb_ld + (arr < 2 ? a.slot_q : a.slot_2) * 128 + (arr & 1) * 64 + c];
        }
        float x[32];                                           // this thread's half of the xq row, kept for the residual

        // ---- xq half row -> registers; AdaLN_q -> split A tiles over the same bytes ----
        tc::mbar_wait(bar_in, it & 1);
        caf_read_half(row_own, sw, x);
        caf_adaln_pair(x, part, tid, gbs, gbs + 64, a.eps, s_ahi, s_alo, r, hf);   // first barrier inside also publishes gbs
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, CAF_NS);
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w), wl = tc::umma_desc_sw128(s_w + 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(tS, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
            }
            tc::umma_commit(bar_mma);
        }
        float p[CAF_MAXJ];                                     // folded score bias of this thread's head, then the probabilities
#pragma unroll
        for (int j = 0; j < CAF_MAXJ; j += 4) {
            const float4 v = ld4(a.sb + (size_t)b * CAF_NS + hf * CAF_KP + j);
            p[j] = v.x; p[j + 1] = v.y; p[j + 2] = v.z; p[j + 3] = v.w;
        }
        tc::mbar_wait(bar_mma, mph & 1);
        ++mph;
        tc::tc_fence_after();

        // ---- softmax of head hf over the clip's NK keys; P (split) -> A tile columns [32 hf, 32 hf + 32) ----
        {
            uint32_t v[32];
            tc::tmem_ld_32x32(tS + lane_sel + hf * CAF_KP, v);     // this head's 32 key slots
            tc::tmem_ld_wait();
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < NK; ++j) { p[j] += __uint_as_float(v[j]); m = fmaxf(m, p[j]); }
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < NK; ++j) { p[j] = expf(p[j] - m); l += p[j]; }
            const float inv = 1.0f / l;
#pragma unroll
            for (int j = 0; j < CAF_MAXJ; ++j) p[j] = j < NK ? p[j] * inv : 0.f;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = p[cc * 8 + i];
                uint4 hh, ll;
                tc::split8(y, hh, ll);
                tc::sts16(s_ahi, r, hf * 4 + cc, hh);
                tc::sts16(s_alo, r, hf * 4 + cc, ll);
            }
            tc::sts16(s_ahi, r, hf * 4 + 3, make_uint4(0, 0, 0, 0));     // key slots 24..31 never hold a key
            tc::sts16(s_alo, r, hf * 4 + 3, make_uint4(0, 0, 0, 0));
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 64);
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w + 16384), wl = tc::umma_desc_sw128(s_w + 24576);
#pragma unroll
            for (int k = 0; k < CAF_NS / 16; ++k) {
                tc::umma_bf16(tO, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                tc::umma_bf16(tO, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                tc::umma_bf16(tO, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
            }
            tc::umma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, mph & 1);
        ++mph;
        tc::tc_fence_after();

        // ---- xq' = (out + bp) + xq -> fp32 boxes (region X, free: both MMAs are complete); AdaLN_2(xq') -> t tiles (region W) ----
        {
            uint32_t v[32];
            tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bpv = ld4(gbs + 4 * 64 + hf * 32 + i);
                x[i] = (__uint_as_float(v[i]) + bpv.x) + x[i]; x[i + 1] = (__uint_as_float(v[i + 1]) + bpv.y) + x[i + 1];
                x[i + 2] = (__uint_as_float(v[i + 2]) + bpv.z) + x[i + 2]; x[i + 3] = (__uint_as_float(v[i + 3]) + bpv.w) + x[i + 3];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            tc::sts16f(row_own + ((j ^ sw) << 4), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
        caf_adaln_pair(x, part, tid, gbs + 128, gbs + 192, a.eps, s_w, s_w + AT_TILE, r, hf);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tma_store_3d(&tm_x, smem, 0, row0, b);
            tc::tma_store_3d(&tm_x, smem + 16384, 32, row0, b);
            tc::tma_store_3d(&tm_thi, smem + CAF_OFF_W, 0, row0, b);
            tc::tma_store_3d(&tm_tlo, smem + CAF_OFF_W + AT_TILE, 0, row0, b);
            tc::tma_store_commit();
        }
    }
    if (tid == 0) tc::tma_store_wait_read<0>();          // smem must outlive the reads; the grid boundary orders the writes
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

template <int NK>
static inline int launch_ca_vertex_fused_t(const CUtensorMap* maps, const CaFusedArgs& a, cudaStream_t st) {
    if (!pmce_configure_smem<ca_vertex_fused_kernel<NK>>(CAF_SMEM)) return 2;
    const int ntiles = a.B * a.qtiles;
    const int cap = 3 * tc_num_sms();
    const int grid = ntiles < cap ? ntiles : cap;
    ca_vertex_fused_kernel<NK><<<grid, CAF_THREADS, CAF_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], a);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// true when the fused kernel covers (heads, keys): the reference's vertex stream has 2 heads; 17 (h36m) / 19 (coco) joints
static inline bool ca_vertex_fused_supported(int heads, int nkeys) { return heads == CAF_H && (nkeys == 17 || nkeys == 19); }

// Per-clip folded operands (made by ca_joint_fold_kernel): kq [B*64, 64], vpt [B*64, 64] split bf16, sb [B, 64] fp32
struct CaFolded {
    __nv_bfloat16 *kq_hi, *kq_lo, *vp_hi, *vp_lo;
    float* sb;
};

// xq [B, N1, 64] fp32 (updated in place), t_hi/t_lo [B, N1, 64] bf16 (AdaLN_2 of the result, split)
static inline int launch_ca_vertex_fused(float* xq, __nv_bfloat16* t_hi, __nv_bfloat16* t_lo, const CaFolded& f, CaFusedArgs a, cudaStream_t st) {
    CUtensorMap maps[7];
    a.qtiles = (a.N1 + 127) / 128;
    a.sb = f.sb;
    if (make_tmap_3d(&maps[0], xq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, a.N1, a.B, 64, 64LL * a.N1, 32, 128, 1) ||
        make_tmap_3d(&maps[1], t_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64, a.N1, a.B, 64, 64LL * a.N1, 64, 128, 1) ||
        make_tmap_3d(&maps[2], t_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64, a.N1, a.B, 64, 64LL * a.N1, 64, 128, 1) ||
        make_tmap_bf16(&maps[3], f.kq_hi, a.B * CAF_NS, 64, 64, CAF_NS) || make_tmap_bf16(&maps[4], f.kq_lo, a.B * CAF_NS, 64, 64, CAF_NS) ||
        make_tmap_bf16(&maps[5], f.vp_hi, a.B * 64, 64, 64, 64) || make_tmap_bf16(&maps[6], f.vp_lo, a.B * 64, 64, 64, 64))
        return 1;
    if (a.N2 == 17) return launch_ca_vertex_fused_t<17>(maps, a, st);
    if (a.N2 == 19) return launch_ca_vertex_fused_t<19>(maps, a, st);
    return 4;
}

// ------------------------------------------------------------------------------------------------------
// Joint side of the vertex cross-attention, one CTA per clip (all of it is 17/19-row work, fp32 on CUDA cores):
//   Jf = W_jp P + b + jpos                       (CoevoDecoder.py:177)      [-> xq_out = Jf + jQ when asked, :182]
//   xk = W_j2v Jf + b + j2v_K                    (:184)
//   K  = Wk AdaLN_k(xk) + bk ;  V = Wv AdaLN_v(Jf) + bv     (:53-55 with :84)
//   KQ = scale K_h Wq_h, sb = scale K_h bq_h, VPt = (V_h Wp[:, h]^T)^T       (the folded operands of ca_vertex_fused_kernel)
// replacing six launches (embed, key projection, 2 x AdaLN, 2 x projection GEMM with 17 live rows per 128-row tile).
// With joints == nullptr the kernel starts from given K / V [B,J,64] (pmce_cross_attn_block on arbitrary key/value streams).
// Thread n (of 64 per row group) keeps row n of a 64x64 weight in registers; activations are broadcast from shared memory.
// ------------------------------------------------------------------------------------------------------
constexpr int JKV_THREADS = 256;
constexpr int JKV_ROWS = CAF_MAXJ;

struct JointFoldArgs {
    const float* joints;                         // [B, J, 3] or nullptr (then K_in / V_in are used)
    const float *K_in, *V_in;                    // [B, J, 64] projected keys / values (only when joints == nullptr)
    const float *wjp, *bjp, *jpos, *jq;          // [64,3], [64], [J,64], [J,64] (jq optional)
    const float *wj2v, *bj2v, *j2vk;             // [64,64], [64], [J,64]
    const float *wk, *bk, *wv, *bv;              // [64,64], [64]
    const float *wq, *bq, *wp;                   // [64,64], [64], [64,64] of the vertex cross-attention
    const float* gb; int gb_ld, slot_k, slot_v;
    float* xq_out;                               // optional [B, J, 64]: Jf + jQ (query stream of the joint branch)
    CaFolded f;
    int J; float eps, scale;
};

constexpr int JKV_WLD = 68;                      // smem row stride (floats) of a staged 64x64 weight: conflict-free float4 rows
constexpr int JKV_WMAT = 64 * JKV_WLD;           // floats per staged weight
constexpr int JKV_SMEM = 4 * JKV_WMAT * 4;       // W_j2v (then Wq) | Wk | Wv | Wp  — 106 KB with the static buffers: two CTAs per SM

// stage one [64,64] fp32 weight into shared memory with cp.async (16-byte chunks, row stride JKV_WLD)
__device__ __forceinline__ void jkv_stage_weight(const float* __restrict__ Wg, float* Ws, int tid) {
    for (int idx = tid; idx < 64 * 16; idx += JKV_THREADS) {
        const int row = idx >> 4, c = (idx & 15) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(Ws + row * JKV_WLD + c)), "l"(Wg + row * 64 + c) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// out[i][n] = W[n,:] . xs[i,:] + bias[n] (+ rowadd[i][n]) for the rows i of this thread's row group; W staged in smem
__device__ __forceinline__ void jkv_matmul(const float* Ws, const float* __restrict__ bias, const float* __restrict__ rowadd,
                                           const float* xs, float* out_s, int J, int n, int rg) {
    float w[64];
#pragma unroll
    for (int k = 0; k < 64; k += 4) { const float4 v = ld4(Ws + n * JKV_WLD + k); w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w; }
    const float b = bias[n];
    for (int i = rg; i < J; i += JKV_THREADS / 64) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            const float4 x = ld4(xs + i * 64 + k);
            a0 = fmaf(w[k], x.x, a0); a1 = fmaf(w[k + 1], x.y, a1);
            a0 = fmaf(w[k + 2], x.z, a0); a1 = fmaf(w[k + 3], x.w, a1);
        }
        float r = (a0 + a1) + b;
        if (rowadd) r += rowadd[i * 64 + n];
        out_s[i * 64 + n] = r;
    }
}

__device__ __forceinline__ void jkv_adaln_row(const float* __restrict__ x, const float* __restrict__ g, float eps, float* __restrict__ y, int lane) {
    const float2 v = *reinterpret_cast<const float2*>(x + lane * 2);
    const float mean = warp_sum(v.x + v.y) * (1.0f / 64.0f);
    const float dx = v.x - mean, dy = v.y - mean;
    const float var = warp_sum(dx * dx + dy * dy) * (1.0f / 63.0f);
    const float inv = 1.0f / (sqrtf(var) + eps);
    const float2 ga = *reinterpret_cast<const float2*>(g + lane * 2), be = *reinterpret_cast<const float2*>(g + 64 + lane * 2);
    y[lane * 2] = ga.x * dx * inv + be.x;
    y[lane * 2 + 1] = ga.y * dy * inv + be.y;
}

struct JointFoldArgs3 { JointFoldArgs blk[3]; };   // blockIdx.y selects the co-evolution block: their joint sides are independent

__global__ void __launch_bounds__(JKV_THREADS)
ca_joint_fold_kernel(const __grid_constant__ JointFoldArgs3 args) {
    const JointFoldArgs& a = args.blk[blockIdx.y];
    constexpr int RS = JKV_ROWS * 64;            // floats per [24][64] buffer
    __shared__ __align__(16) float buf[6 * RS];
    float *Jf = buf, *Xk = buf + RS, *Nk = buf + 2 * RS, *Nv = buf + 3 * RS, *Ks = buf + 4 * RS, *Vs = buf + 5 * RS;
    float* VPs = buf;                            // [64][65] staging of VPt^T, aliases Jf/Xk/Nk once K and V are final
    static_assert(CAF_NS * 65 <= 3 * RS, "VPs must not reach Ks/Vs");
    extern __shared__ __align__(16) float wsm[];   // the five 64x64 weights, all in flight from the first instruction
    float *Wj2v = wsm, *Wk = wsm + JKV_WMAT, *Wv = wsm + 2 * JKV_WMAT, *Wp = wsm + 3 * JKV_WMAT, *Wq = wsm;   // Wq reuses W_j2v's slot
    const int b = blockIdx.x, tid = threadIdx.x, J = a.J;
    const int n = tid & 63, rg = tid >> 6;
    if (a.joints) {
        jkv_stage_weight(a.wj2v, Wj2v, tid); cp_async_commit();
        jkv_stage_weight(a.wk, Wk, tid); jkv_stage_weight(a.wv, Wv, tid); jkv_stage_weight(a.wp, Wp, tid); cp_async_commit();
    } else {
        jkv_stage_weight(a.wq, Wq, tid); jkv_stage_weight(a.wp, Wp, tid); cp_async_commit();
    }
    if (a.joints) {
        const float* P = a.joints + (size_t)b * J * 3;
        {
            const float w0 = a.wjp[n * 3], w1 = a.wjp[n * 3 + 1], w2 = a.wjp[n * 3 + 2], bb = a.bjp[n];
            for (int i = rg; i < J; i += JKV_THREADS / 64) {
                const float f = ((w0 * P[i * 3] + w1 * P[i * 3 + 1]) + w2 * P[i * 3 + 2]) + bb + a.jpos[i * 64 + n];
                Jf[i * 64 + n] = f;
                if (a.xq_out) a.xq_out[((size_t)b * J + i) * 64 + n] = f + a.jq[i * 64 + n];
            }
        }
        cp_async_wait<1>();
        __syncthreads();
        jkv_matmul(Wj2v, a.bj2v, a.j2vk, Jf, Xk, J, n, rg);
        __syncthreads();
        jkv_stage_weight(a.wq, Wq, tid); cp_async_commit();      // W_j2v's slot is free: Wq lands under the LayerNorm / K / V phases
        {
            const int warp = tid >> 5, lane = tid & 31;
            const float* g = a.gb + (size_t)b * a.gb_ld;
            for (int i = warp; i < J; i += JKV_THREADS / 32) {
                jkv_adaln_row(Xk + i * 64, g + a.slot_k * 128, a.eps, Nk + i * 64, lane);
                jkv_adaln_row(Jf + i * 64, g + a.slot_v * 128, a.eps, Nv + i * 64, lane);
            }
        }
        cp_async_wait<1>();
        __syncthreads();
        jkv_matmul(Wk, a.bk, nullptr, Nk, Ks, J, n, rg);
        jkv_matmul(Wv, a.bv, nullptr, Nv, Vs, J, n, rg);
    } else {
        for (int idx = tid; idx < J * 16; idx += JKV_THREADS) {
            st4(Ks + idx * 4, ld4(a.K_in + (size_t)b * J * 64 + idx * 4));
            st4(Vs + idx * 4, ld4(a.V_in + (size_t)b * J * 64 + idx * 4));
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- fold: KQ[32h+j][c] = scale sum_d K[j][32h+d] Wq[32h+d][c]  (thread = output channel c, coalesced weight columns) ----
#pragma unroll 1
    for (int h = 0; h < CAF_H; ++h) {
        float wc[32];
#pragma unroll
        for (int d = 0; d < 32; ++d) wc[d] = Wq[(h * 32 + d) * JKV_WLD + n];
        for (int j = rg; j < CAF_KP; j += JKV_THREADS / 64) {
            float acc = 0.f;
            if (j < J) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int d = 0; d < 32; d += 4) {
                    const float4 k = ld4(Ks + j * 64 + h * 32 + d);
                    a0 = fmaf(k.x, wc[d], a0); a1 = fmaf(k.y, wc[d + 1], a1);
                    a0 = fmaf(k.z, wc[d + 2], a0); a1 = fmaf(k.w, wc[d + 3], a1);
                }
                acc = (a0 + a1) * a.scale;
            }
            __nv_bfloat16 hi, lo;
            tc::split_bf16(acc, hi, lo);
            const size_t o = ((size_t)b * CAF_NS + h * CAF_KP + j) * 64 + n;
            a.f.kq_hi[o] = hi; a.f.kq_lo[o] = lo;
        }
    }
    if (tid < CAF_NS) {                          // sb[32h+j] = scale sum_d bq[32h+d] K[j][32h+d]
        const int h = tid / CAF_KP, j = tid % CAF_KP;
        float acc = 0.f;
        if (j < J)
            for (int d = 0; d < 32; ++d) acc = fmaf(a.bq[h * 32 + d], Ks[j * 64 + h * 32 + d], acc);
        a.f.sb[(size_t)b * CAF_NS + tid] = acc * a.scale;
    }
    // ---- fold: VPt[n][32h+j] = sum_d V[j][32h+d] Wp[n][32h+d]  (thread = output channel n, weight row in registers) ----
    {
        float w[64];
#pragma unroll
        for (int k = 0; k < 64; k += 4) { const float4 v = ld4(Wp + n * JKV_WLD + k); w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w; }
        for (int idx = rg; idx < CAF_NS; idx += JKV_THREADS / 64) {
            const int h = idx / CAF_KP, j = idx % CAF_KP;
            float acc = 0.f;
            if (j < J) {
                float a0 = 0.f, a1 = 0.f;
                if (h == 0) {
#pragma unroll
                    for (int d = 0; d < 32; d += 4) {
                        const float4 v = ld4(Vs + j * 64 + d);
                        a0 = fmaf(v.x, w[d], a0); a1 = fmaf(v.y, w[d + 1], a1); a0 = fmaf(v.z, w[d + 2], a0); a1 = fmaf(v.w, w[d + 3], a1);
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < 32; d += 4) {
                        const float4 v = ld4(Vs + j * 64 + 32 + d);
                        a0 = fmaf(v.x, w[32 + d], a0); a1 = fmaf(v.y, w[32 + d + 1], a1); a0 = fmaf(v.z, w[32 + d + 2], a0); a1 = fmaf(v.w, w[32 + d + 3], a1);
                    }
                }
                acc = a0 + a1;
            }
            VPs[idx * 65 + n] = acc;            // Jf/Xk/Nk (aliased) were last read before the barrier that published Ks/Vs
        }
    }
    __syncthreads();
    for (int e = tid; e < 64 * 32; e += JKV_THREADS) {       // VPt rows [64][64], split-bf16, coalesced
        const int row = e >> 5, c = (e & 31) * 2;
        const float v0 = VPs[c * 65 + row], v1 = VPs[(c + 1) * 65 + row];
        uint32_t h2, l2;
        tc::split_bf16x2(v0, v1, h2, l2);
        const size_t o = ((size_t)b * 64 + row) * 64 + c;
        *reinterpret_cast<uint32_t*>(a.f.vp_hi + o) = h2;
        *reinterpret_cast<uint32_t*>(a.f.vp_lo + o) = l2;
    }
}
