// Fused vertex<-joint CrossAttention of a CrossAttentionBlock (CoevoDecoder.py:47-62 inside :82-85), one kernel per block:
//     xq' = xq + Wp . MHA( Wq AdaLN_q(xq) + bq ,  K ,  V ) + bp            (K, V: the 17/19 projected joint rows of the clip)
// replacing AdaLN -> Wq GEMM -> attention -> Wp GEMM(+residual) (four launches, each a round trip of the [B*431, 64] stream
// through HBM) with one pass: the stream is read once and written once.  (AdaLN_2 + Mlp of the block: mlp_fused.cuh.)
//
// Few keys => everything that is per clip is folded into two per-clip operands by ca_joint_fold_kernel (below):
//   n        = (x - mean) / (std_unbiased + eps)                  per row, the only LayerNorm work left per element (1 FMA)
//   S        = n KQ'^T + sb'      KQ'[24h+j][c] = log2e scale (K_h Wq_h)[j][c] gamma_q[c]
//                                 sb' [24h+j]   = log2e (scale K_hj . bq_h + sum_c scale (K_h Wq_h)[j][c] beta_q[c])
//   P_h      = 2^(S_h - max_j S_h) / sum                                   (softmax over the clip's NK keys, per head)
//   xq'      = xq + P VP'         VP'[24h+j][n] = (V_h Wp[:, h]^T)[j][n] + bp[n] / 2     (rows of P_h sum to 1: 2 heads carry bp)
// so a 128-row item is: TMA load -> row statistics -> one FMA per element -> split-bf16 A tiles -> tcgen05 128x48x64 ->
// softmax of 2 x NK scores in registers -> P tiles -> tcgen05 128x64x48 ON TOP of xq (preloaded into the accumulator) -> TMA store.
//
// The per-item chain (load, two MMA round trips, store drain) is ~7 us of latency against ~1 us of issue work, so throughput
// comes from items in flight: ONE CTA per SM with CA_G = 4 independent groups of 4 warps, each walking its own items:
//   * thread r of a group owns tile row r = TMEM lane r; the whole 64-wide row passes through its registers once, so the
//     LayerNorm statistics are thread-local (no cross-thread reduction);
//   * ONE 32 KB buffer per group changes roles in place - the two [128][32 fp32] TMA boxes of x -> the A tiles hi|lo (row r of
//     a tile occupies exactly the bytes of row r of a box: no hazard between threads, no barrier) -> the P tiles -> the fp32
//     output boxes for the TMA store; the residual x does not stay in registers either: it is written into the output
//     accumulator (tcgen05.st) and the second MMA accumulates onto it;
//   * the operands are compact (48 key slots: KQ' [48][64] K-major, VP' [48][64] MN-major, 24 KB with hi|lo), per group;
//   * no producer warps: the group's first thread issues the group's TMA loads at the moment it learns a buffer is free
//     (KQ' after the first MMA's commit, VP' after the second's, x after the store has drained) and its MMAs;
//   * three group-wide named barriers per item (A tiles written, P tiles written, output staged).
// Shared memory: 4 x (32 + 24) KB = 224 KB; TMEM: 4 x (64 + 64) columns.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"

constexpr int CAF_KP = 24;                              // key slots per head (num_joint <= 24: 17 h36m, 19 coco; the rest are zero)
constexpr int CAF_MAXJ = 24;
constexpr int CAF_H = 2;                                // heads of the vertex stream (CoevoDecoder.py:140)
constexpr int CAF_NS = CAF_H * CAF_KP;                  // 48 score columns / key slots: slot 24 h + j = (head h, key j)
constexpr int CA_G = 4;                                 // groups = items in flight per CTA
constexpr int CA_THREADS = CA_G * 128;
constexpr int CA_X = 2 * 128 * 128;                     // x boxes = A tiles hi|lo = P tiles = output boxes (32 KB, in place)
constexpr int CA_WT = CAF_NS * 128;                     // one operand tile [48][64 bf16]: 6 KB
constexpr int CA_W = 4 * CA_WT;                         // KQ' hi | KQ' lo | VP' hi | VP' lo
constexpr int CA_GBUF = CA_X + CA_W;                    // 56 KB per group (1024-byte aligned pieces)
constexpr int CA_OFF_BAR = CA_G * CA_GBUF;
constexpr int CA_OFF_WE = CA_OFF_BAR + 256;             // W_e [64,3] of the embed mode (768 B)
constexpr int CA_SMEM = CA_OFF_WE + 768 + 1024;         // barriers + W_e + alignment slack
static_assert(CA_SMEM <= 227 * 1024, "shared memory");
static_assert(CA_WT % 1024 == 0, "operand tiles must stay 1024-byte aligned");

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
// TMEM -> registers: this warp's 32 lanes (rows) x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// registers -> TMEM: this warp's 32 lanes (rows) x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
          "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
          "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace tc

struct CaFusedArgs {
    const float* sb;         // [B, 48] folded score bias sb' (log2 domain)
    int B, N1, N2, qtiles;
    float eps;
    // embed mode (coords != nullptr): the query stream does not exist in memory yet - row i of clip b is
    //   xq[b,i,:] = W_e coords[b,i,:] + E[i,:],  E = b_e + pos + Q_embed  (CoevoDecoder.py:178,182; E is one [N1,64] table per block)
    // the kernel loads E tiles (shared by every clip: L2 hits) where it would load xq, adds the 3-term product per element and
    // writes the updated stream to xq: the block's first pass over the vertex stream costs a write instead of write + read + write
    const float* coords;     // [B, N1, 3] or nullptr (then xq is read and updated in place)
    const float* we;         // W_e [64,3] (device): staged once per CTA in shared memory, read as 16-byte broadcasts
};

template <int NK>
__global__ void __launch_bounds__(CA_THREADS, 1)
ca_vertex_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_kq_hi,
                       const __grid_constant__ CUtensorMap tm_kq_lo, const __grid_constant__ CUtensorMap tm_vp_hi,
                       const __grid_constant__ CUtensorMap tm_vp_lo, const __grid_constant__ CUtensorMap tm_e,
                       CaFusedArgs a) {
    static_assert(NK <= CAF_MAXJ, "too many keys");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CA_OFF_BAR);
    uint64_t* in_full = bars;                    // [G] TMA bytes of the x tile
    uint64_t* kq_full = bars + CA_G;             // [G] TMA bytes of KQ' hi|lo
    uint64_t* vp_full = bars + 2 * CA_G;         // [G] TMA bytes of VP' hi|lo
    uint64_t* mma_bar = bars + 3 * CA_G;         // [G] tcgen05.commit: S complete / O complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4 * CA_G);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = warp >> 2, r = tid & 127;
    const bool leader = r == 0;
    const int ntiles = a.B * a.qtiles;
    uint8_t* gptr = smem + g * CA_GBUF;
    const int tile0 = blockIdx.x + g * gridDim.x, tstep = CA_G * gridDim.x;       // this group's items
    const bool embed = a.coords != nullptr;
    float* s_we = reinterpret_cast<float*>(smem + CA_OFF_WE);
    if (embed && tid < 192) s_we[tid] = a.we[tid];                                 // visible after the __syncthreads below
    if (leader) {
        tc::mbar_init(&in_full[g], 1); tc::mbar_init(&kq_full[g], 1); tc::mbar_init(&vp_full[g], 1); tc::mbar_init(&mma_bar[g], 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, CA_G * 128);
    if (leader) {
        pdl_wait();                 // x and the folded operands are the previous kernels' outputs
        if (tile0 < ntiles) {       // the group's first item is requested before the CTA assembles
            const int b = tile0 / a.qtiles, row0 = (tile0 % a.qtiles) * 128;
            tc::mbar_arrive_expect_tx(&in_full[g], CA_X);
            tc::tma_load_3d(gptr, embed ? &tm_e : &tm_x, &in_full[g], 0, row0, embed ? 0 : b);
            tc::tma_load_3d(gptr + 16384, embed ? &tm_e : &tm_x, &in_full[g], 32, row0, embed ? 0 : b);
            tc::mbar_arrive_expect_tx(&kq_full[g], 2 * CA_WT);
            tc::tma_load_2d(gptr + CA_X, &tm_kq_hi, &kq_full[g], 0, b * CAF_NS);
            tc::tma_load_2d(gptr + CA_X + CA_WT, &tm_kq_lo, &kq_full[g], 0, b * CAF_NS);
            tc::mbar_arrive_expect_tx(&vp_full[g], 2 * CA_WT);
            tc::tma_load_2d(gptr + CA_X + 2 * CA_WT, &tm_vp_hi, &vp_full[g], 0, b * CAF_NS);
            tc::tma_load_2d(gptr + CA_X + 3 * CA_WT, &tm_vp_lo, &vp_full[g], 0, b * CAF_NS);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();
    pdl_wait();

    const uint32_t s_x = sbase + g * CA_GBUF, s_ahi = s_x, s_alo = s_x + AT_TILE, s_w = s_x + CA_X;
    const uint32_t tS = tmem_base + g * 128, tO = tS + 64;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const int sw = r & 7;
    const uint32_t row_addr = s_x + r * 128;
    uint32_t j = 0;
    for (int tile = tile0; tile < ntiles; tile += tstep, ++j) {
        const int b = tile / a.qtiles, row0 = (tile % a.qtiles) * 128;
        const int next = tile + tstep;
        const bool has_next = next < ntiles;
        const int nb = has_next ? next / a.qtiles : 0, nrow0 = has_next ? (next % a.qtiles) * 128 : 0;

        // ---- the row: shared memory -> registers -> (a) the output accumulator (residual), (b) statistics, (c) A tiles in place ----
        float p0 = 0.f, p1 = 0.f, p2 = 0.f;
        if (embed && row0 + r < a.N1) {           // this row's coordinates: in flight while the tile arrives
            const float* cp = a.coords + ((size_t)b * a.N1 + row0 + r) * 3;
            p0 = __ldg(cp); p1 = __ldg(cp + 1); p2 = __ldg(cp + 2);
        }
        tc::mbar_wait(&in_full[g], j & 1);
        float inv, nmi;
        {
            float x[64];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 v = tc::lds16(row_addr + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4));
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            if (embed) {
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {          // 4 channels = 12 weights = three 16-byte broadcasts
                    const float4 wa = *reinterpret_cast<const float4*>(s_we + 12 * c4), wb = *reinterpret_cast<const float4*>(s_we + 12 * c4 + 4),
                                 wc = *reinterpret_cast<const float4*>(s_we + 12 * c4 + 8);
                    x[4 * c4] += (wa.x * p0 + wa.y * p1) + wa.z * p2;
                    x[4 * c4 + 1] += (wa.w * p0 + wb.x * p1) + wb.y * p2;
                    x[4 * c4 + 2] += (wb.z * p0 + wb.w * p1) + wc.x * p2;
                    x[4 * c4 + 3] += (wc.y * p0 + wc.z * p1) + wc.w * p2;
                }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(x[hf * 32 + i]);
                tc::tmem_st_32x32(tO + lane_sel + hf * 32, v);
            }
            // AdaLayerNorm statistics (CoevoDecoder.py:23-29: unbiased std, eps added to the std), thread-local
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
            inv = 1.0f / (sqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / 63.0f)) + a.eps);
            nmi = -mean * inv;
            // n = (x - mean) inv -> split-bf16 A tiles over the SAME bytes (row r of the hi / lo tile = row r of box 0 / 1)
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
        }
        tc::tmem_st_wait();
        tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::bar_sync_group(1 + g);
        if (leader) {
            tc::mbar_wait(&kq_full[g], j & 1);
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, CAF_NS);
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w), wl = tc::umma_desc_sw128(s_w + CA_WT);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(tS, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                tc::umma_bf16(tS, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
            }
            tc::umma_commit(&mma_bar[g]);
        }
        // folded score bias of both heads (the same 2 x NK floats for every row of the clip: L1/L2 broadcast), in flight while the MMA runs
        constexpr int NKV = (NK + 3) / 4;
        float4 sbv[2][NKV];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < NKV; ++i) sbv[h][i] = __ldg(reinterpret_cast<const float4*>(a.sb + (size_t)b * CAF_NS + h * CAF_KP) + i);
        tc::mbar_wait(&mma_bar[g], 0);
        tc::tc_fence_after();
        if (leader && has_next) {                 // KQ' is free: the next item's tiles arrive under the rest of this item
            tc::mbar_arrive_expect_tx(&kq_full[g], 2 * CA_WT);
            tc::tma_load_2d(gptr + CA_X, &tm_kq_hi, &kq_full[g], 0, nb * CAF_NS);
            tc::tma_load_2d(gptr + CA_X + CA_WT, &tm_kq_lo, &kq_full[g], 0, nb * CAF_NS);
        }

        // ---- softmax of each head over the clip's NK keys (log2 domain); P (split) -> A tile columns [24 h, 24 h + 24) ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float p[CAF_KP];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint32_t v[8];
                tc::tmem_ld_32x8(tS + lane_sel + h * CAF_KP + q * 8, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) p[q * 8 + i] = __uint_as_float(v[i]);
            }
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < NK; ++i) {
                const float4 t = sbv[h][i >> 2];
                p[i] += (i & 3) == 0 ? t.x : ((i & 3) == 1 ? t.y : ((i & 3) == 2 ? t.z : t.w));
                m = fmaxf(m, p[i]);
            }
            float l = 0.f;
#pragma unroll
            for (int i = 0; i < NK; ++i) { p[i] = tc::ex2_approx(p[i] - m); l += p[i]; }
            const float il = 1.0f / l;
#pragma unroll
            for (int i = 0; i < CAF_KP; ++i) p[i] = i < NK ? p[i] * il : 0.f;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = p[cc * 8 + i];
                uint4 hh, ll;
                tc::split8(y, hh, ll);
                tc::sts16(s_ahi, r, h * 3 + cc, hh);
                tc::sts16(s_alo, r, h * 3 + cc, ll);
            }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::bar_sync_group(1 + g);
        if (leader) {
            tc::mbar_wait(&vp_full[g], j & 1);
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32_bmn(128, 64);
            const uint64_t ah = tc::umma_desc_sw128(s_ahi), al = tc::umma_desc_sw128(s_alo);
            const uint64_t wh = tc::umma_desc_sw128(s_w + 2 * CA_WT), wl = tc::umma_desc_sw128(s_w + 3 * CA_WT);
#pragma unroll
            for (int k = 0; k < CAF_NS / 16; ++k) {          // every MMA ACCUMULATES: the accumulator already holds xq
                const uint64_t bh = wh + (uint64_t)(k * 2048 >> 4), bl = wl + (uint64_t)(k * 2048 >> 4);     // 16 key slots = two 8-row groups
                tc::umma_bf16(tO, tc::umma_desc_advance_k(al, k), bh, idesc, 1);
                tc::umma_bf16(tO, tc::umma_desc_advance_k(ah, k), bl, idesc, 1);
                tc::umma_bf16(tO, tc::umma_desc_advance_k(ah, k), bh, idesc, 1);
            }
            tc::umma_commit(&mma_bar[g]);
        }
        tc::mbar_wait(&mma_bar[g], 1);
        tc::tc_fence_after();
        if (leader && has_next) {                 // VP' is free
            tc::mbar_arrive_expect_tx(&vp_full[g], 2 * CA_WT);
            tc::tma_load_2d(gptr + CA_X + 2 * CA_WT, &tm_vp_hi, &vp_full[g], 0, nb * CAF_NS);
            tc::tma_load_2d(gptr + CA_X + 3 * CA_WT, &tm_vp_lo, &vp_full[g], 0, nb * CAF_NS);
        }

        // ---- xq' = accumulator -> fp32 boxes over the same buffer (both MMAs have read it) -> TMA store ----
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c)
                tc::sts16f(row_addr + hf * 16384 + ((c ^ sw) << 4), make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                                                   __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])));
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::bar_sync_group(1 + g);
        if (leader) {
            tc::tma_store_3d(&tm_x, gptr, 0, row0, b);
            tc::tma_store_3d(&tm_x, gptr + 16384, 32, row0, b);
            tc::tma_store_commit();
            tc::tma_store_wait_read<0>();         // the buffer has been read: the next item's rows may land in it
            if (has_next) {
                tc::mbar_arrive_expect_tx(&in_full[g], CA_X);
                tc::tma_load_3d(gptr, embed ? &tm_e : &tm_x, &in_full[g], 0, nrow0, embed ? 0 : nb);
                tc::tma_load_3d(gptr + 16384, embed ? &tm_e : &tm_x, &in_full[g], 32, nrow0, embed ? 0 : nb);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, CA_G * 128);
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
    if (!pmce_configure_smem<ca_vertex_fused_kernel<NK>>(CA_SMEM)) return 2;
    const int ntiles = a.B * a.qtiles;
    const int cap = tc_num_sms();                      // one CTA per SM (225 KB of shared memory), CA_G items in flight each
    const int grid = ntiles < cap ? ntiles : cap;
    return pmce_launch(ca_vertex_fused_kernel<NK>, dim3(grid), dim3(CA_THREADS), CA_SMEM, st, 0, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], a)
                   == cudaSuccess ? 0 : 3;
}

// true when the fused kernel covers (heads, keys): the reference's vertex stream has 2 heads; 17 (h36m) / 19 (coco) joints
static inline bool ca_vertex_fused_supported(int heads, int nkeys) { return heads == CAF_H && (nkeys == 17 || nkeys == 19); }

// Per-clip folded operands (made by ca_joint_fold_kernel): kq [B*48, 64] (row = key slot, K-major over the channel c),
// vp [B*48, 64] (row = key slot, MN-major over the output channel n), split bf16; sb [B, 48] fp32
struct CaFolded {
    __nv_bfloat16 *kq_hi, *kq_lo, *vp_hi, *vp_lo;
    float* sb;
};

// Embed mode of the kernel: coords [B, N1, 3], table E [N1, 64] (made by ca_embed_table_kernel), we = W_e [64, 3]; all device memory
struct CaEmbed {
    const float* coords;
    const float* table;
    const float* we;         // [64, 3] device
};

// E[i][c] = (b_e[c] + pos[i][c]) + Q[i][c] for up to three blocks in one launch (grid.y = block)
struct CaEmbedTableArgs {
    const float *bias[3], *pos[3], *q[3];
    float* out[3];
    int n;                   // rows (N1)
};
__global__ void ca_embed_table_kernel(CaEmbedTableArgs a) {
    const int k = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.n * 64) return;
    a.out[k][idx] = (a.bias[k][idx & 63] + a.pos[k][idx]) + a.q[k][idx];
}

// xq [B, N1, 64] fp32: updated in place, or (embed != nullptr) written from the coordinates
static inline int launch_ca_vertex_fused(float* xq, const CaFolded& f, CaFusedArgs a, cudaStream_t st, const CaEmbed* embed = nullptr) {
    CUtensorMap maps[6];
    a.qtiles = (a.N1 + 127) / 128;
    a.sb = f.sb;
    a.coords = nullptr;
    if (embed) {
        a.coords = embed->coords;
        a.we = embed->we;
        if (make_tmap_3d(&maps[5], embed->table, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, a.N1, 1, 64, 64LL * a.N1, 32, 128, 1)) return 1;
    }
    if (make_tmap_3d(&maps[0], xq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 64, a.N1, a.B, 64, 64LL * a.N1, 32, 128, 1) ||
        make_tmap_bf16(&maps[1], f.kq_hi, a.B * CAF_NS, 64, 64, CAF_NS) || make_tmap_bf16(&maps[2], f.kq_lo, a.B * CAF_NS, 64, 64, CAF_NS) ||
        make_tmap_bf16(&maps[3], f.vp_hi, a.B * CAF_NS, 64, 64, CAF_NS) || make_tmap_bf16(&maps[4], f.vp_lo, a.B * CAF_NS, 64, 64, CAF_NS))
        return 1;
    if (!embed) maps[5] = maps[0];
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
//   VP' = V_h Wp[:, h]^T + bp / 2 on the live key slots     (the folded operands of ca_vertex_fused_kernel: AdaLN_q's
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
    pdl_enter();          // the weights (constants) are already in flight; joints / K / V / gamma-beta are other kernels' outputs
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
    // ---- fold: KQ[24h+j][c] = scale sum_d K[j][32h+d] Wq[32h+d][c]  (thread = output channel c, coalesced weight columns) ----
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
    if (tid < CAF_NS) {                          // sb'[24h+j] = log2e (scale sum_d bq[32h+d] K[j][32h+d] + sum_c KQ[32h+j][c] beta_q[c])
        const int h = tid / CAF_KP, j = tid % CAF_KP;
        float acc = 0.f;
        if (j < J)
            for (int d = 0; d < 32; ++d) acc = fmaf(a.bq[h * 32 + d], Ks[j * 64 + h * 32 + d], acc);
        a.f.sb[(size_t)b * CAF_NS + tid] = (acc * a.scale + sbacc[tid]) * LOG2E;
    }
    // ---- fold: VP'[24h+j][n] = sum_d V[j][32h+d] Wp[n][32h+d] + bp[n]/2  (thread = output channel n, weight row in registers) ----
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
            __nv_bfloat16 hi, lo;                // VP' row = key slot, the 64 threads of a row group write 64 consecutive channels
            tc::split_bf16(acc, hi, lo);
            const size_t o = ((size_t)b * CAF_NS + idx) * 64 + n;
            a.f.vp_hi[o] = hi; a.f.vp_lo[o] = lo;
        }
    }
}
