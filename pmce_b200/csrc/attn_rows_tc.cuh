// Tensor-core attention for head_dim 32 and up to 448 keys (the co-evolution decoder's vertex self-attention, 431 x 431,
// CoevoDecoder.py:119-131): the WHOLE score row of a 128-query tile lives in TMEM, so the softmax is exact two-pass (no
// online rescaling between key chunks) and the probabilities never touch shared memory:
//   item   = (sequence b, head h, 128-query tile)
//   S      = (Q scale log2e) K^T  for all key chunks at once: tcgen05 128 x 128 (x 64 for the last chunk), split-bf16 with the
//            hi|lo halves CONCATENATED into one 128-byte swizzled row (head_dim 32 = 64-byte operand rows):
//            A = [Q_hi|Q_lo], B1 = [K_hi|K_hi] (k16 steps 0-3: Q_hi K_hi + Q_lo K_hi), B2 = [K_lo| - ] (steps 0-1: Q_hi K_lo)
//            -> 448 fp32 TMEM columns
//   m      = row max (each of the row's two threads scans two key chunks; the halves meet through shared memory)
//   P      = 2^(S - m), written back IN PLACE over S as packed bf16 pairs (tcgen05.st): per 32-key group 16 columns P_hi then
//            16 columns P_lo - the A operand of the next MMA is read straight from TMEM (tcgen05.mma with A in TMEM)
//   D      = P_hi [V_hi|V_lo] + P_lo [V_hi|V_lo], accumulated over the chunks in TMEM (64 columns; V is an MN-major B operand,
//            consumed in place); o = (D[:, :32] + D[:, 32:]) / sum P   -> split-bf16 -> global
// Operands arrive by TMA, already in the layout the MMAs read: the block's qkv projection (gemm_tc.cuh, epilogue TC_ATTN32)
// writes one bf16 tensor [rows, 512] of 128-byte (row, head) records QC | K1 | K2 | VC.
// Warp roles (one CTA per SM, 576 threads):
//   warps 0-15   math: thread (row r, quarter q = warp / 4) owns TMEM lane r and key chunk q (max / sum meet through shared memory)
//   warp 16      TMA producer (one lane): Q + K tiles are re-requested as soon as the S MMAs of the current item have completed,
//                V after its last P V MMA - the next item's operands land while the current softmax runs
//   warp 17      MMA issuer (+ TMEM owner): S for all chunks in one batch, then P V chunk by chunk as the quarters publish them
// Hand-offs: mbarriers for TMA and tcgen05.commit completions; named barriers for math -> issuer.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"
#include "mlp_fused.cuh"

constexpr int AR_MAXK = 448;                       // key slots: 3 chunks of 128 + one of 64
constexpr int AR_OFF_Q = 0;                        // [Q_hi|Q_lo]            [128][128 B]
constexpr int AR_OFF_K1 = AT_TILE;                 // [K_hi|K_hi] x 4 chunks
constexpr int AR_OFF_K2 = AR_OFF_K1 + 4 * AT_TILE; // [K_lo| -  ] x 4 chunks
constexpr int AR_OFF_V = AR_OFF_K2 + 4 * AT_TILE;  // [V_hi|V_lo] x 4 chunks
constexpr int AR_OFF_MX = AR_OFF_V + 4 * AT_TILE;  // row max | row sum exchange between the quarters: 2 x [4][128] fp32
constexpr int AR_OFF_BAR = AR_OFF_MX + 4096;
constexpr int AR_SMEM = AR_OFF_BAR + 128 + 1024;
constexpr int AR_THREADS = 18 * 32;
static_assert(AR_SMEM <= 227 * 1024, "shared memory");
// named barriers
constexpr int AR_BAR_MATH = 1;                     // the 512 math threads
constexpr int AR_BAR_P0 = 2;                       // ids 2..5: P of chunk c written (128 math threads arrive, issuer warp syncs)
constexpr int AR_BAR_EPI = 6;                      // 512 math threads arrive (output accumulator read), issuer syncs

namespace tc {
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (128 lanes x K/2 packed bf16 pairs per column) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
}  // namespace tc

__global__ void __launch_bounds__(AR_THREADS, 1)
attn_rows_tc_kernel(const __grid_constant__ CUtensorMap tm_att, SplitOut Os, AttnAddr ao, int N1, int N2, int nseq, int H) {
    constexpr int D = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    float* mx = reinterpret_cast<float*>(smem + AR_OFF_MX);                 // [4][128] row max of each quarter
    float* ls = mx + 512;                                                   // [4][128] row sum of each quarter
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AR_OFF_BAR);
    uint64_t* kq_full = bars;          // TMA bytes: Q, K1, K2 tiles
    uint64_t* v_full = bars + 1;       // TMA bytes: V tiles
    uint64_t* s_done = bars + 2;       // tcgen05.commit: all S MMAs complete (scores ready; Q/K tiles free)
    uint64_t* o_done = bars + 3;       // tcgen05.commit: all P V MMAs complete (output ready; V tiles and the S/P columns free)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tm_att);
        tc::mbar_init(kq_full, 1); tc::mbar_init(v_full, 1); tc::mbar_init(s_done, 1); tc::mbar_init(o_done, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 17) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tS = tmem_base, tD = tmem_base + AR_MAXK;
    pdl_trigger();
    pdl_wait();          // the prologue above overlapped the previous kernel; its outputs are visible from here

    const int qtiles = (N1 + 127) / 128;
    const int nchunks = (N2 + 127) / 128;                                   // <= 4 (N2 <= 448)
    const int nwork = nseq * H * qtiles;          // work -> (seq, head, qtile), qtile fastest (K/V of a (seq, head) stay L2-hot)

    if (warp == 16) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++it) {
                const int qt = work % qtiles, sh = work / qtiles;
                const int h = sh % H, s = sh / H;
                if (it > 0) tc::mbar_wait(s_done, (it - 1) & 1);             // the previous item's S MMAs have read the Q / K tiles
                tc::mbar_arrive_expect_tx(kq_full, (uint32_t)((1 + 2 * nchunks) * AT_TILE));
                tc::tma_load_3d(smem + AR_OFF_Q, &tm_att, kq_full, h * 64, qt * 128, s);
                for (int c = 0; c < nchunks; ++c) {
                    tc::tma_load_3d(smem + AR_OFF_K1 + c * AT_TILE, &tm_att, kq_full, 128 + h * 64, c * 128, s);
                    tc::tma_load_3d(smem + AR_OFF_K2 + c * AT_TILE, &tm_att, kq_full, 256 + h * 64, c * 128, s);
                }
                if (it > 0) tc::mbar_wait(o_done, (it - 1) & 1);             // the previous item's P V MMAs have read the V tiles
                tc::mbar_arrive_expect_tx(v_full, (uint32_t)(nchunks * AT_TILE));
                for (int c = 0; c < nchunks; ++c) tc::tma_load_3d(smem + AR_OFF_V + c * AT_TILE, &tm_att, v_full, 384 + h * 64, c * 128, s);
            }
        }
    } else if (warp == 17) {
        // ================= MMA issuer =================
        uint32_t it = 0;
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++it) {
            tc::mbar_wait(kq_full, it & 1);
            if (lane == 0) {
                tc::tc_fence_after();
                const uint64_t qc = tc::umma_desc_sw128(sb + AR_OFF_Q);
                for (int c = 0; c < nchunks; ++c) {
                    const uint32_t idesc = c < 3 ? tc::umma_idesc_bf16_f32(128, 128) : tc::umma_idesc_bf16_f32(128, 64);
                    const uint64_t khh = tc::umma_desc_sw128(sb + AR_OFF_K1 + c * AT_TILE), kl = tc::umma_desc_sw128(sb + AR_OFF_K2 + c * AT_TILE);
                    // Q_hi K_lo (steps 0,1 of [K_lo|K_lo]) first, then [Q_hi|Q_lo] [K_hi|K_hi]
                    tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, 0), tc::umma_desc_advance_k(kl, 0), idesc, 0);
                    tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, 1), tc::umma_desc_advance_k(kl, 1), idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, k), tc::umma_desc_advance_k(khh, k), idesc, 1);
                }
                tc::umma_commit(s_done);
            }
            __syncwarp();
            tc::mbar_wait(v_full, it & 1);
            tc::bar_sync_n(AR_BAR_EPI, 512 + 32);                            // the previous item's output accumulator has been read
            for (int c = 0; c < nchunks; ++c) {
                tc::bar_sync_n(AR_BAR_P0 + c, 128 + 32);                     // P of chunk c is in TMEM
                if (lane == 0) {
                    tc::tc_fence_after();
                    constexpr uint32_t idesc = tc::umma_idesc_bf16_f32_bmn(128, 64);
                    const uint64_t vc = tc::umma_desc_sw128(sb + AR_OFF_V + c * AT_TILE);
                    const int ksteps = c < 3 ? 8 : 4;
                    for (int t = 0; t < ksteps; ++t) {
                        const uint32_t pa = tS + c * 128 + 32 * (t >> 1) + 8 * (t & 1);       // P_hi of keys [16 t, 16 t + 16); P_lo 16 columns on
                        const uint64_t bv = vc + (uint64_t)(t * 2048 >> 4);                   // 16 keys = two 8-row groups
                        tc::umma_bf16_ts(tD, pa + 16, bv, idesc, (c | t) != 0);
                        tc::umma_bf16_ts(tD, pa, bv, idesc, 1);
                    }
                    if (c == nchunks - 1) tc::umma_commit(o_done);
                }
                __syncwarp();
            }
        }
        tc::bar_sync_n(AR_BAR_EPI, 512 + 32);       // pairs with the math threads' arrive after their LAST output read
    } else {
        // ================= math: thread = (row r, quarter q4); quarter q4 owns key chunk q4 =================
        const int r = tid & 127, q4 = warp >> 2;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        const bool have = q4 < nchunks;
        const int nk = N2 - q4 * 128;                                        // live keys of this quarter's chunk (when it has one)
        const int slots = q4 < 3 ? 128 : 64;
        uint32_t it = 0;
        tc::bar_arrive_n(AR_BAR_EPI, 512 + 32);                              // "no previous output to read" for the first item
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++it) {
            const int qt = work % qtiles, sh = work / qtiles;
            const int h = sh % H, s = sh / H;
            const int qrow = qt * 128 + r;
            tc::mbar_wait(s_done, it & 1);
            tc::tc_fence_after();
            // ---- pass 1: row max over this quarter's chunk ----
            float m = -INFINITY;
            if (have) {
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    if (q * 32 >= nk) break;
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tS + lane_sel + q4 * 128 + q * 32, v);
                    tc::tmem_ld_wait();
                    if (q * 32 + 32 <= nk) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (q * 32 + i < nk) m = fmaxf(m, __uint_as_float(v[i]));
                    }
                }
            }
            mx[q4 * 128 + r] = m;
            tc::bar_sync_n(AR_BAR_MATH, 512);
            m = fmaxf(fmaxf(mx[r], mx[128 + r]), fmaxf(mx[256 + r], mx[384 + r]));       // (-inf from a quarter without keys is harmless)
            // ---- pass 2: P = 2^(S - m) in place over S (bf16 hi / lo pairs), row sum ----
            float l = 0.f;
            if (have) {
#pragma unroll 1
                for (int q = 0; q * 32 < slots; ++q) {
                    uint32_t v[32], o[32];
                    tc::tmem_ld_32x32(tS + lane_sel + q4 * 128 + q * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = (q * 32 + i < nk) ? tc::ex2_approx(__uint_as_float(v[i]) - m) : 0.f;
                        const float p1 = (q * 32 + i + 1 < nk) ? tc::ex2_approx(__uint_as_float(v[i + 1]) - m) : 0.f;
                        l += p0 + p1;
                        tc::split_bf16x2(p0, p1, o[i >> 1], o[16 + (i >> 1)]);      // columns [0,16): P_hi pairs, [16,32): P_lo pairs
                    }
                    tc::tmem_st_32x32(tS + lane_sel + q4 * 128 + q * 32, o);
                }
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::bar_arrive_n(AR_BAR_P0 + q4, 128 + 32);
            }
            ls[q4 * 128 + r] = l;
            tc::bar_sync_n(AR_BAR_MATH, 512);
            l = (ls[r] + ls[128 + r]) + (ls[256 + r] + ls[384 + r]);
            // ---- output: this quarter's 8 of the 32 dims ----
            tc::mbar_wait(o_done, it & 1);
            tc::tc_fence_after();
            {
                uint32_t d0[8], d1[8];
                tc::tmem_ld_32x8(tD + lane_sel + q4 * 8, d0);
                tc::tmem_ld_32x8(tD + lane_sel + 32 + q4 * 8, d1);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::bar_arrive_n(AR_BAR_EPI, 512 + 32);
                if (qrow < N1) {
                    const float inv = 1.0f / l;
                    const size_t ob = (size_t)(ao.seq(s) + (long long)qrow * ao.tok) * ao.ld + h * D + q4 * 8;
#pragma unroll
                    for (int i = 0; i < 8; i += 4)
                        store_split4(Os, ob + i, make_float4((__uint_as_float(d0[i]) + __uint_as_float(d1[i])) * inv,
                                                             (__uint_as_float(d0[i + 1]) + __uint_as_float(d1[i + 1])) * inv,
                                                             (__uint_as_float(d0[i + 2]) + __uint_as_float(d1[i + 2])) * inv,
                                                             (__uint_as_float(d0[i + 3]) + __uint_as_float(d1[i + 3])) * inv));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 17) tc::tmem_dealloc(tmem_base, 512);
}

// nseq sequences of N tokens (N <= 448), H = 2 heads of 32: `att` is the [nseq * N, 512] bf16 record tensor written by the qkv
// projection's TC_ATTN32 epilogue (q carries softmax scale * log2e); split-bf16 output (addr ao).
static inline bool attn_rows_tc_supported(int heads, int N) { return heads == 2 && N >= 1 && N <= AR_MAXK; }
static inline float attn_rows_qscale() { return 1.4426950408889634f / sqrtf(32.0f); }

static inline int launch_attn_rows_tc(const __nv_bfloat16* att, SplitOut Os, AttnAddr ao, int nseq, int H, int N, cudaStream_t st) {
    if (!pmce_configure_smem<attn_rows_tc_kernel>(AR_SMEM)) return 2;
    CUtensorMap tm;
    if (make_tmap_bf16_3d(&tm, att, TC_ATT_LD, N, nseq, TC_ATT_LD, (long long)N * TC_ATT_LD, 128, 1)) return 1;
    const long long work = (long long)nseq * H * ((N + 127) / 128);
    const int sms = tc_num_sms();
    const int grid = (int)(work < sms ? (work < 1 ? 1 : work) : sms);
    if (pmce_launch(attn_rows_tc_kernel, dim3(grid), dim3(AR_THREADS), AR_SMEM, st, 0, tm, Os, ao, N, N, nseq, H) != cudaSuccess) return 3;
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
