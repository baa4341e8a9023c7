// Fused vertex<-joint CrossAttention of a CrossAttentionBlock (CoevoDecoder.py:47-62 inside :82-85), one kernel per block:
//     xq' = xq + Wp . MHA( Wq AdaLN_q(xq) + bq ,  K ,  V ) + bp            (K, V: the 17/19 projected joint rows of the clip)
// replacing AdaLN -> Wq GEMM -> attention -> Wp GEMM(+residual) (four launches, each a round trip of the [B*431, 64] stream
// through HBM) with one pass: the stream is read once and written once.  (AdaLN_2 + Mlp of the block: mlp_fused.cuh.)
//
// Few keys => everything that is per clip is folded into two per-clip operands by ca_joint_fold_kernel (below):
//   n        = (x - mean) / (std_unbiased + eps)                  per row, the only LayerNorm work left per element (1 FMA)
//   S        = n KQ'^T + sb'      KQ'[32h+j][c] = log2e scale (K_h Wq_h)[j][c] gamma_q[c]
//                                 sb' [32h+j]   = log2e (scale K_hj . bq_h + sum_c scale (K_h Wq_h)[j][c] beta_q[c])
//   P_h      = 2^(S_h - max_j S_h) / sum                                   (softmax over the clip's NK keys, per head)
//   xq'      = xq + P VPt'^T      VPt'[n][32h+j] = (V_h Wp[:, h]^T)[j][n] + bp[n] / 2     (rows of P_h sum to 1: 2 heads carry bp)
// so a 128-row item is: TMA load -> row statistics -> one FMA per element -> split-bf16 A tiles -> tcgen05 128x64x64 ->
// softmax of 2 x NK scores in registers -> P tiles -> tcgen05 128x64x64 -> + xq -> TMA store.
//
// Warp-specialised, persistent, ONE CTA per SM with CA2_G independent consumer groups (items in flight):
//   warps 0-3 / 4-7   consumer group 0 / 1: thread r owns tile row r = TMEM lane r with the WHOLE 64-wide row in registers, so
//                     the LayerNorm statistics are thread-local (no cross-thread reduction, no barrier); 4 group-wide named
//                     barriers per item (A buffer free, A tiles written, P tiles written, output staged); the group's first
//                     thread issues its MMAs and its TMA store.
//   warp 8            producer of the x tiles (one lane): the group's IN buffer is released as soon as its 128 rows are in
//                     registers, so the NEXT item's rows stream in under the whole chain of the current one.
//   warp 9 / 10       producers of the per-clip operand tiles KQ' / VPt' (16 KB each, L2 hits for 3 of a clip's 4 tiles): KQ' is
//                     released by the commit of the first MMA and VPt' by the second, so the next item's KQ' is resident long
//                     before its rows are normalised.
// Shared memory per group: IN 32 KB (two [128][32 fp32] SW128 boxes) | A 32 KB (A tiles hi|lo, then P hi|lo, then the fp32
// output boxes for the TMA store) | W 32 KB (KQ' hi|lo, VPt' hi|lo) = 96 KB; TMEM 128 columns per group (S | O).
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"

constexpr int CAF_KP = 32;                              // key slots per head (num_joint <= 24 live: 17 h36m, 19 coco; the rest are zero)
constexpr int CAF_MAXJ = 24;
constexpr int CAF_H = 2;                                // heads of the vertex stream (CoevoDecoder.py:140)
constexpr int CAF_NS = CAF_H * CAF_KP;                  // 64 score columns: column 32 h + j = (head h, key j)
constexpr int CA2_G = 2;                                // consumer groups = items in flight per CTA
constexpr int CA2_THREADS = CA2_G * 128 + 96;           // + x-tile producer warp + two operand producer warps
constexpr int CA2_IN = 2 * 128 * 128;                   // two [128][32 fp32] boxes
constexpr int CA2_A = 2 * AT_TILE;                      // A tiles hi | lo ([128][64 bf16] each) = the two fp32 output boxes later
constexpr int CA2_W = 4 * 8192;                         // KQ' hi | KQ' lo | VPt' hi | VPt' lo ([64][64 bf16] each)
constexpr int CA2_GBUF = CA2_IN + CA2_A + CA2_W;        // 96 KB per group
constexpr int CA2_OFF_BAR = CA2_G * CA2_GBUF;
constexpr int CA2_SMEM = CA2_OFF_BAR + 256 + 1024;      // barriers + alignment slack

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
__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, rel. error 2^-22.5 (MUFU.EX2)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
}  // namespace tc

struct CaFusedArgs {
    const float* sb;         // [B, 64] folded score bias sb' (log2 domain)
    float* xq;               // [B, N1, 64] the stream itself (rows are stored straight from registers)
    int B, N1, N2, qtiles;
    float eps;
    int tma_out;             // 1: stage the output rows in shared memory and TMA-store them (A/B knob PMCE_CA_TMA_OUT)
};

template <int NK>
__global__ void __launch_bounds__(CA2_THREADS, 1)
ca_vertex_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_kq_hi,
                       const __grid_constant__ CUtensorMap tm_kq_lo, const __grid_constant__ CUtensorMap tm_vp_hi,
                       const __grid_constant__ CUtensorMap tm_vp_lo, CaFusedArgs a) {
    static_assert(NK <= CAF_MAXJ, "too many keys");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CA2_OFF_BAR);
    uint64_t* in_full = bars;                    // [G] TMA bytes of the x tile
    uint64_t* in_empty = bars + CA2_G;           // [G] 4 arrivals: every consumer warp has its rows in registers
    uint64_t* kq_full = bars + 2 * CA2_G;        // [G] TMA bytes of KQ' hi|lo
    uint64_t* kq_empty = bars + 3 * CA2_G;       // [G] tcgen05.commit after the item's first MMA
    uint64_t* vp_full = bars + 4 * CA2_G;        // [G] TMA bytes of VPt' hi|lo
    uint64_t* vp_empty = bars + 5 * CA2_G;       // [G] tcgen05.commit after the item's second MMA
    uint64_t* mma_bar = bars + 6 * CA2_G;        // [G] tcgen05.commit: S complete / O complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 7 * CA2_G);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = a.B * a.qtiles;
    if (tid == 0) {
        for (int g = 0; g < CA2_G; ++g) {
            tc::mbar_init(&in_full[g], 1); tc::mbar_init(&in_empty[g], 4);
            tc::mbar_init(&kq_full[g], 1); tc::mbar_init(&kq_empty[g], 1); tc::mbar_init(&vp_full[g], 1); tc::mbar_init(&vp_empty[g], 1);
            tc::mbar_init(&mma_bar[g], 1);
        }
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        // the first item of every group is requested here, before TMEM is allocated and the CTA assembles: at small batches
        // (one or two items per CTA) this DRAM round trip is a third of the kernel
        for (int g = 0; g < CA2_G; ++g) {
            const int tile = blockIdx.x + g * gridDim.x;
            if (tile >= ntiles) break;
            const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;
            uint8_t* gbuf = smem + g * CA2_GBUF;
            tc::mbar_arrive_expect_tx(&in_full[g], CA2_IN);
            tc::tma_load_3d(gbuf, &tm_x, &in_full[g], 0, row0, b);
            tc::tma_load_3d(gbuf + 16384, &tm_x, &in_full[g], 32, row0, b);
            tc::mbar_arrive_expect_tx(&kq_full[g], CA2_W / 2);
            tc::tma_load_2d(gbuf + CA2_IN + CA2_A, &tm_kq_hi, &kq_full[g], 0, b * CAF_NS);
            tc::tma_load_2d(gbuf + CA2_IN + CA2_A + 8192, &tm_kq_lo, &kq_full[g], 0, b * CAF_NS);
            tc::mbar_arrive_expect_tx(&vp_full[g], CA2_W / 2);
            tc::tma_load_2d(gbuf + CA2_IN + CA2_A + 16384, &tm_vp_hi, &vp_full[g], 0, b * 64);
            tc::tma_load_2d(gbuf + CA2_IN + CA2_A + 24576, &tm_vp_lo, &vp_full[g], 0, b * 64);
        }
    }
    if (warp == CA2_G * 4) tc::tmem_alloc(tmem_ptr_smem, CA2_G * 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == CA2_G * 4) {
        // ================= producer of the x tiles: item n of this CTA -> group n % G =================
        if (lane == 0) {
            uint32_t n = CA2_G;                         // (the first round was requested in the prologue)
            for (int tile = blockIdx.x + CA2_G * gridDim.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int g = n % CA2_G;
                const uint32_t j = n / CA2_G;
                const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;
                uint8_t* in = smem + g * CA2_GBUF;
                tc::mbar_wait(&in_empty[g], (j & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&in_full[g], CA2_IN);
                tc::tma_load_3d(in, &tm_x, &in_full[g], 0, row0, b);
                tc::tma_load_3d(in + 16384, &tm_x, &in_full[g], 32, row0, b);
            }
        }
    } else if (warp == CA2_G * 4 + 1) {
        // ================= producer of the KQ' tiles =================
        if (lane == 0) {
            uint32_t n = CA2_G;
            for (int tile = blockIdx.x + CA2_G * gridDim.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int g = n % CA2_G;
                const uint32_t j = n / CA2_G;
                const int b = tile / a.qtiles;
                uint8_t* w = smem + g * CA2_GBUF + CA2_IN + CA2_A;
                tc::mbar_wait(&kq_empty[g], (j & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&kq_full[g], CA2_W / 2);
                tc::tma_load_2d(w, &tm_kq_hi, &kq_full[g], 0, b * CAF_NS);
                tc::tma_load_2d(w + 8192, &tm_kq_lo, &kq_full[g], 0, b * CAF_NS);
            }
        }
    } else if (warp == CA2_G * 4 + 2) {
        // ================= producer of the VPt' tiles =================
        if (lane == 0) {
            uint32_t n = CA2_G;
            for (int tile = blockIdx.x + CA2_G * gridDim.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int g = n % CA2_G;
                const uint32_t j = n / CA2_G;
                const int b = tile / a.qtiles;
                uint8_t* w = smem + g * CA2_GBUF + CA2_IN + CA2_A + CA2_W / 2;
                tc::mbar_wait(&vp_empty[g], (j & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&vp_full[g], CA2_W / 2);
                tc::tma_load_2d(w, &tm_vp_hi, &vp_full[g], 0, b * 64);
                tc::tma_load_2d(w + 8192, &tm_vp_lo, &vp_full[g], 0, b * 64);
            }
        }
    } else {
        // ================= consumers: group g = warp / 4, thread r owns tile row r =================
        const int g = warp >> 2, r = tid & 127;
        const uint32_t gbuf = sbase + g * CA2_GBUF;
        const uint32_t s_in = gbuf, s_ahi = gbuf + CA2_IN, s_alo = s_ahi + AT_TILE, s_w = gbuf + CA2_IN + CA2_A;
        const uint32_t tS = tmem_base + g * 128, tO = tS + 64;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        const int sw = r & 7;
        const uint32_t in_row = s_in + r * 128, out_row = s_ahi + r * 128;
        const bool leader = r == 0;
        bool store_pending = false;
        uint32_t j = 0;
        for (int tile = blockIdx.x + g * gridDim.x; tile < ntiles; tile += CA2_G * gridDim.x, ++j) {
            const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;

            // ---- the row -> registers (kept for the residual); the IN buffer goes back to the producer at once ----
            float x[64];
            tc::mbar_wait(&in_full[g], j & 1);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 v = tc::lds16(in_row + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4));
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&in_empty[g]);

            // ---- AdaLayerNorm statistics (CoevoDecoder.py:23-29: unbiased std, eps added to the std), thread-local ----
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 64; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
            const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / 64.0f);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                const float d0 = x[i] - mean, d1 = x[i + 1] - mean, d2 = x[i + 2] - mean, d3 = x[i + 3] - mean;
                q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
            }
            const float inv = 1.0f / (sqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / 63.0f)) + a.eps);
            const float nmi = -mean * inv;

            // ---- the A buffer is free once the previous item's TMA store has read it ----
            if (a.tma_out) {
                if (leader && store_pending) tc::tma_store_wait_read<0>();
                tc::bar_sync_group(1 + g);
            }

            // ---- n = (x - mean) inv -> split-bf16 A tiles ----
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = fmaf(x[8 * c + i], inv, nmi);
                uint4 hh, ll;
                tc::split8(y, hh, ll);
                tc::sts16(s_ahi, r, c, hh);
                tc::sts16(s_alo, r, c, ll);
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::bar_sync_group(1 + g);
            if (leader) {
                tc::mbar_wait(&kq_full[g], j & 1);
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
                tc::umma_commit(&mma_bar[g]);
                tc::umma_commit(&kq_empty[g]);                      // KQ' goes back to its producer: the next item's tiles arrive early
            }
            // folded score bias of both heads (the same 2 x NK floats for every row of the clip: L1/L2 broadcast), in flight
            // while the MMA runs
            constexpr int NKV = (NK + 3) / 4;                      // float4 loads per head
            float4 sbv[2][NKV];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < NKV; ++i) sbv[h][i] = __ldg(reinterpret_cast<const float4*>(a.sb + (size_t)b * CAF_NS + h * CAF_KP) + i);
            tc::mbar_wait(&mma_bar[g], 0);
            tc::tc_fence_after();

            // ---- softmax of each head over the clip's NK keys (log2 domain); P (split) -> A tile columns [32 h, 32 h + 32) ----
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tc::tmem_ld_32x32(tS + lane_sel + h * CAF_KP, v);
                tc::tmem_ld_wait();
                float p[CAF_MAXJ];
                float m = -INFINITY;
#pragma unroll
                for (int i = 0; i < NK; ++i) {
                    const float4 t = sbv[h][i >> 2];
                    const float bias = (i & 3) == 0 ? t.x : ((i & 3) == 1 ? t.y : ((i & 3) == 2 ? t.z : t.w));
                    p[i] = __uint_as_float(v[i]) + bias;
                    m = fmaxf(m, p[i]);
                }
                float l = 0.f;
#pragma unroll
                for (int i = 0; i < NK; ++i) { p[i] = tc::ex2_approx(p[i] - m); l += p[i]; }
                const float il = 1.0f / l;
#pragma unroll
                for (int i = 0; i < CAF_MAXJ; ++i) p[i] = i < NK ? p[i] * il : 0.f;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = p[cc * 8 + i];
                    uint4 hh, ll;
                    tc::split8(y, hh, ll);
                    tc::sts16(s_ahi, r, h * 4 + cc, hh);
                    tc::sts16(s_alo, r, h * 4 + cc, ll);
                }
                tc::sts16(s_ahi, r, h * 4 + 3, make_uint4(0, 0, 0, 0));     // key slots 24..31 never hold a key
                tc::sts16(s_alo, r, h * 4 + 3, make_uint4(0, 0, 0, 0));
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::bar_sync_group(1 + g);
            if (leader) {
                tc::mbar_wait(&vp_full[g], j & 1);
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
                tc::umma_commit(&mma_bar[g]);
                tc::umma_commit(&vp_empty[g]);
            }
            tc::mbar_wait(&mma_bar[g], 1);
            tc::tc_fence_after();

            // ---- xq' = xq + P VPt'^T -> fp32 boxes in the A buffer (both MMAs have read it) -> TMA store ----
            if (a.tma_out) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        tc::sts16f(out_row + hf * 16384 + ((c ^ sw) << 4),
                                   make_float4(x[hf * 32 + 4 * c] + __uint_as_float(v[4 * c]), x[hf * 32 + 4 * c + 1] + __uint_as_float(v[4 * c + 1]),
                                               x[hf * 32 + 4 * c + 2] + __uint_as_float(v[4 * c + 2]), x[hf * 32 + 4 * c + 3] + __uint_as_float(v[4 * c + 3])));
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::bar_sync_group(1 + g);
                if (leader) {
                    tc::tma_store_3d(&tm_x, smem + g * CA2_GBUF + CA2_IN, 0, row0, b);
                    tc::tma_store_3d(&tm_x, smem + g * CA2_GBUF + CA2_IN + 16384, 32, row0, b);
                    tc::tma_store_commit();
                    store_pending = true;
                }
            } else {
                // the thread's 256-byte row goes straight to global memory: no staging, no proxy fence, no group barrier, and the
                // next item's A tiles never wait for a TMA store to drain the buffer
                float* orow = a.xq + ((size_t)b * a.N1 + row0 + r) * 64;
                const bool live = row0 + r < a.N1;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v);
                    tc::tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            __stcs(reinterpret_cast<float4*>(orow + hf * 32 + 4 * c),
                                   make_float4(x[hf * 32 + 4 * c] + __uint_as_float(v[4 * c]), x[hf * 32 + 4 * c + 1] + __uint_as_float(v[4 * c + 1]),
                                               x[hf * 32 + 4 * c + 2] + __uint_as_float(v[4 * c + 2]), x[hf * 32 + 4 * c + 3] + __uint_as_float(v[4 * c + 3])));
                    }
                }
                tc::tc_fence_before();       // orders the TMEM reads above before the group barrier that precedes the next item's MMA
            }
        }
        if (leader && store_pending) tc::tma_store_wait_read<0>();     // smem must outlive the reads; the grid boundary orders the writes
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == CA2_G * 4) tc::tmem_dealloc(tmem_base, CA2_G * 128);
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
    if (!pmce_configure_smem<ca_vertex_fused_kernel<NK>>(CA2_SMEM)) return 2;
    const int ntiles = a.B * a.qtiles;
    const int cap = tc_num_sms();                      // one CTA per SM (193 KB of shared memory), CA2_G items in flight each
    const int grid = ntiles < cap ? ntiles : cap;
    ca_vertex_fused_kernel<NK><<<grid, CA2_THREADS, CA2_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], a);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// true when the fused kernel covers (heads, keys): the reference's vertex stream has 2 heads; 17 (h36m) / 19 (coco) joints
static inline bool ca_vertex_fused_supported(int heads, int nkeys) { return heads == CAF_H && (nkeys == 17 || nkeys == 19); }

// Per-clip folded operands (made by ca_joint_fold_kernel): kq [B*64, 64], vpt [B*64, 64] split bf16, sb [B, 64] fp32
struct CaFolded {
    __nv_bfloat16 *kq_hi, *kq_lo, *vp_hi, *vp_lo;
    float* sb;
};

// xq [B, N1, 64] fp32, updated in place
static inline int launch_ca_vertex_fused(float* xq, const CaFolded& f, CaFusedArgs a, cudaStream_t st) {
    CUtensorMap maps[5];
    a.qtiles = (a.N1 + 127) / 128;
    a.sb = f.sb;
    a.xq = xq;
    static int tma_out = -1;
    if (tma_out < 0) tma_out = pmce_env_int("PMCE_CA_TMA_OUT", 1) ? 1 : 0;   // measured: direct row stores 27.6 us vs 23.8 us staged + TMA (B=256)
    a.tma_out = tma_out;
    if (make_tmap_3d(&maps[0], xq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, a.N1, a.B, 64, 64LL * a.N1, 32, 128, 1) ||
        make_tmap_bf16(&maps[1], f.kq_hi, a.B * CAF_NS, 64, 64, CAF_NS) || make_tmap_bf16(&maps[2], f.kq_lo, a.B * CAF_NS, 64, 64, CAF_NS) ||
        make_tmap_bf16(&maps[3], f.vp_hi, a.B * 64, 64, 64, 64) || make_tmap_bf16(&maps[4], f.vp_lo, a.B * 64, 64, 64, 64))
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
//   KQ' = log2e scale (K_h Wq_h) diag(gamma_q), sb' = log2e (scale K_h bq_h + scale (K_h Wq_h) beta_q),
//   VPt' = (V_h Wp[:, h]^T)^T + bp / 2 on the live key slots     (the folded operands of ca_vertex_fused_kernel: AdaLN_q's
//   per-clip gamma / beta, the softmax's log2e and the output bias all live in them)
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
    const float *wq, *bq, *wp, *bp;              // [64,64], [64], [64,64], [64] of the vertex cross-attention
    const float* gb; int gb_ld, slot_k, slot_v, slot_q;
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
    __shared__ float sbacc[CAF_NS];              // sum_c KQ[s][c] beta_q[c]: two warps (c = 0..31, 32..63) add into each slot
    if (tid < CAF_NS) sbacc[tid] = 0.f;
    cp_async_wait<0>();
    __syncthreads();
    constexpr float LOG2E = 1.4426950408889634f;
    const float gq = a.gb[(size_t)b * a.gb_ld + a.slot_q * 128 + n], bq_ = a.gb[(size_t)b * a.gb_ld + a.slot_q * 128 + 64 + n];
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
            // (a + b == b + a: the two warps' contributions commute, the sum is deterministic)
            const float part = warp_sum(acc * bq_);
            if ((tid & 31) == 0 && j < J) atomicAdd(&sbacc[h * CAF_KP + j], part);
            acc *= gq * LOG2E;
            __nv_bfloat16 hi, lo;
            tc::split_bf16(acc, hi, lo);
            const size_t o = ((size_t)b * CAF_NS + h * CAF_KP + j) * 64 + n;
            a.f.kq_hi[o] = hi; a.f.kq_lo[o] = lo;
        }
    }
    __syncthreads();                             // sbacc complete
    if (tid < CAF_NS) {                          // sb'[32h+j] = log2e (scale sum_d bq[32h+d] K[j][32h+d] + sum_c KQ[32h+j][c] beta_q[c])
        const int h = tid / CAF_KP, j = tid % CAF_KP;
        float acc = 0.f;
        if (j < J)
            for (int d = 0; d < 32; ++d) acc = fmaf(a.bq[h * 32 + d], Ks[j * 64 + h * 32 + d], acc);
        a.f.sb[(size_t)b * CAF_NS + tid] = (acc * a.scale + sbacc[tid]) * LOG2E;
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
                acc = (a0 + a1) + 0.5f * a.bp[n];           // every softmax row sums to 1: the two heads carry the output bias
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
