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
// Warp roles (one CTA per SM, 544 threads):
//   warps 0-7    math: thread (row r, half h = warp / 4) owns TMEM lane r; half h handles key chunks h and h + 2
//   warps 8-15   loaders: fp32 q/k/v rows -> split-bf16 -> swizzled operand tiles. K/Q tiles are released as soon as the S MMAs
//                have completed and V after the last P V MMA, so the next item's operands are converted while the current
//                item's softmax runs
//   warp 16      MMA issuer (+ TMEM owner): S for all chunks in one batch, then P V chunk by chunk as the math halves publish them
// Hand-offs: mbarriers for loader <-> issuer and tcgen05.commit completions; named barriers for math -> issuer.
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
constexpr int AR_OFF_MX = AR_OFF_V + 4 * AT_TILE;  // row max / row sum exchange between the halves: 2 x 128 fp32
constexpr int AR_OFF_BAR = AR_OFF_MX + 1024;
constexpr int AR_SMEM = AR_OFF_BAR + 128 + 1024;
constexpr int AR_THREADS = 17 * 32;
static_assert(AR_SMEM <= 227 * 1024, "shared memory");
// named barriers
constexpr int AR_BAR_MATH = 1;                     // the 256 math threads
constexpr int AR_BAR_P0 = 2;                       // ids 2..5: P of chunk c written (128 math threads arrive, issuer warp syncs)
constexpr int AR_BAR_EPI = 6;                      // 256 math threads arrive (output accumulator read), issuer syncs

namespace tc {
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
attn_rows_tc_kernel(const float* __restrict__ Q, AttnAddr aq, const float* __restrict__ K, const float* __restrict__ V, AttnAddr akv, SplitOut Os,
                    AttnAddr ao, int N1, int N2, int nseq, int H, float qscale) {
    constexpr int D = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    float* mx = reinterpret_cast<float*>(smem + AR_OFF_MX);                 // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AR_OFF_BAR);
    uint64_t* kq_full = bars;          // 256 loader arrivals: Q, K1, K2 tiles written
    uint64_t* v_full = bars + 1;       // 256 loader arrivals: V tiles written
    uint64_t* s_done = bars + 2;       // tcgen05.commit: all S MMAs complete (scores ready; Q/K tiles free)
    uint64_t* o_done = bars + 3;       // tcgen05.commit: all P V MMAs complete (output ready; V tiles and the S/P columns free)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(kq_full, 256); tc::mbar_init(v_full, 256); tc::mbar_init(s_done, 1); tc::mbar_init(o_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 16) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tS = tmem_base, tD = tmem_base + AR_MAXK;

    const int qtiles = (N1 + 127) / 128;
    const int nchunks = (N2 + 127) / 128;                                   // <= 4 (N2 <= 448)
    const int nwork = nseq * H * qtiles;          // work -> (seq, head, qtile), qtile fastest (K/V of a (seq, head) stay L2-hot)

    if (warp >= 8 && warp < 16) {
        // ================= loaders: thread = (tile row, half of the 32-wide head slice) =================
        const int lt = tid - 256;
        const int lr = lt & 127, hf = lt >> 7;
        uint32_t it = 0;
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++it) {
            const int qt = work % qtiles, sh = work / qtiles;
            const int h = sh % H, s = sh / H;
            // Q row (scaled) + K rows of every chunk -> registers first (the global loads fly while the previous item's S runs)
            float4 qx[4];
            {
                const int qrow = qt * 128 + lr;
                if (qrow < N1) {
                    const size_t qb = (size_t)(aq.seq(s) + (long long)qrow * aq.tok) * aq.ld + h * D + hf * 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) qx[i] = ld4(Q + qb + i * 4);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) qx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float4 kx[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int key = c * 128 + lr;
                if (c < nchunks && key < N2) {
                    const size_t kb = (size_t)(akv.seq(s) + (long long)key * akv.tok) * akv.ld + h * D + hf * 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) kx[c][i] = ld4(K + kb + i * 4);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) kx[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (it > 0) tc::mbar_wait(s_done, (it - 1) & 1);                 // the previous item's S MMAs have read the Q / K tiles
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float x[8];
                uint4 hh, ll;
                x[0] = qx[2 * cc].x * qscale; x[1] = qx[2 * cc].y * qscale; x[2] = qx[2 * cc].z * qscale; x[3] = qx[2 * cc].w * qscale;
                x[4] = qx[2 * cc + 1].x * qscale; x[5] = qx[2 * cc + 1].y * qscale; x[6] = qx[2 * cc + 1].z * qscale; x[7] = qx[2 * cc + 1].w * qscale;
                tc::split8(x, hh, ll);
                tc::sts16(sb + AR_OFF_Q, lr, hf * 2 + cc, hh); tc::sts16(sb + AR_OFF_Q, lr, 4 + hf * 2 + cc, ll);            // [Q_hi | Q_lo]
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c >= nchunks || (c == 3 && lr >= 64)) continue;          // (the fourth chunk holds 64 key slots)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float x[8];
                    uint4 hh, ll;
                    x[0] = kx[c][2 * cc].x; x[1] = kx[c][2 * cc].y; x[2] = kx[c][2 * cc].z; x[3] = kx[c][2 * cc].w;
                    x[4] = kx[c][2 * cc + 1].x; x[5] = kx[c][2 * cc + 1].y; x[6] = kx[c][2 * cc + 1].z; x[7] = kx[c][2 * cc + 1].w;
                    tc::split8(x, hh, ll);
                    tc::sts16(sb + AR_OFF_K1 + c * AT_TILE, lr, hf * 2 + cc, hh); tc::sts16(sb + AR_OFF_K1 + c * AT_TILE, lr, 4 + hf * 2 + cc, hh);   // [K_hi | K_hi]
                    tc::sts16(sb + AR_OFF_K2 + c * AT_TILE, lr, hf * 2 + cc, ll);                                                                      // [K_lo |  -  ]
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(kq_full);
            // V rows
            float4 vx[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int key = c * 128 + lr;
                if (c < nchunks && key < N2) {
                    const size_t kb = (size_t)(akv.seq(s) + (long long)key * akv.tok) * akv.ld + h * D + hf * 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) vx[c][i] = ld4(V + kb + i * 4);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) vx[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (it > 0) tc::mbar_wait(o_done, (it - 1) & 1);                 // the previous item's P V MMAs have read the V tiles
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c >= nchunks || (c == 3 && lr >= 64)) continue;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float x[8];
                    uint4 hh, ll;
                    x[0] = vx[c][2 * cc].x; x[1] = vx[c][2 * cc].y; x[2] = vx[c][2 * cc].z; x[3] = vx[c][2 * cc].w;
                    x[4] = vx[c][2 * cc + 1].x; x[5] = vx[c][2 * cc + 1].y; x[6] = vx[c][2 * cc + 1].z; x[7] = vx[c][2 * cc + 1].w;
                    tc::split8(x, hh, ll);
                    tc::sts16(sb + AR_OFF_V + c * AT_TILE, lr, hf * 2 + cc, hh); tc::sts16(sb + AR_OFF_V + c * AT_TILE, lr, 4 + hf * 2 + cc, ll);     // [V_hi | V_lo]
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(v_full);
        }
    } else if (warp == 16) {
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
                    // Q_hi K_lo (steps 0,1 of [K_lo|-]) first, then [Q_hi|Q_lo] [K_hi|K_hi]
                    tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, 0), tc::umma_desc_advance_k(kl, 0), idesc, 0);
                    tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, 1), tc::umma_desc_advance_k(kl, 1), idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_bf16(tS + c * 128, tc::umma_desc_advance_k(qc, k), tc::umma_desc_advance_k(khh, k), idesc, 1);
                }
                tc::umma_commit(s_done);
            }
            __syncwarp();
            tc::mbar_wait(v_full, it & 1);
            tc::bar_sync_n(AR_BAR_EPI, 256 + 32);                            // the previous item's output accumulator has been read
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
    } else if (warp < 8) {
        // ================= math: thread = (row r, half h); half h owns key chunks h and h + 2 =================
        const int r = tid & 127, h2 = warp >> 2;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t it = 0;
        tc::bar_arrive_n(AR_BAR_EPI, 256 + 32);                              // "no previous output to read" for the first item
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++it) {
            const int qt = work % qtiles, sh = work / qtiles;
            const int h = sh % H, s = sh / H;
            const int qrow = qt * 128 + r;
            tc::mbar_wait(s_done, it & 1);
            tc::tc_fence_after();
            // ---- pass 1: row max over this half's chunks ----
            float m = -INFINITY;
#pragma unroll 1
            for (int c = h2; c < nchunks; c += 2) {
                const int nk = N2 - c * 128;                                 // live keys of the chunk (> 0)
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    if (q * 32 >= nk) break;
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tS + lane_sel + c * 128 + q * 32, v);
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
            mx[h2 * 128 + r] = m;
            tc::bar_sync_n(AR_BAR_MATH, 256);
            m = fmaxf(m, mx[(h2 ^ 1) * 128 + r]);                            // (-inf from a half without chunks is harmless)
            // ---- pass 2: P = 2^(S - m) in place over S (bf16 hi / lo pairs), row sum ----
            float l = 0.f;
#pragma unroll 1
            for (int c = h2; c < nchunks; c += 2) {
                const int nk = N2 - c * 128;
                const int slots = c < 3 ? 128 : 64;
#pragma unroll 1
                for (int q = 0; q * 32 < slots; ++q) {
                    uint32_t v[32], o[32];
                    tc::tmem_ld_32x32(tS + lane_sel + c * 128 + q * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = (q * 32 + i < nk) ? tc::ex2_approx(__uint_as_float(v[i]) - m) : 0.f;
                        const float p1 = (q * 32 + i + 1 < nk) ? tc::ex2_approx(__uint_as_float(v[i + 1]) - m) : 0.f;
                        l += p0 + p1;
                        tc::split_bf16x2(p0, p1, o[i >> 1], o[16 + (i >> 1)]);      // columns [0,16): P_hi pairs, [16,32): P_lo pairs
                    }
                    tc::tmem_st_32x32(tS + lane_sel + c * 128 + q * 32, o);
                }
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::bar_arrive_n(AR_BAR_P0 + c, 128 + 32);
            }
            __syncwarp();
            tc::bar_sync_n(AR_BAR_MATH, 256);                                // everyone has read the partner's max: mx can carry the sums
            mx[h2 * 128 + r] = l;
            tc::bar_sync_n(AR_BAR_MATH, 256);
            l += mx[(h2 ^ 1) * 128 + r];
            // ---- output: this half's 16 of the 32 dims ----
            tc::mbar_wait(o_done, it & 1);
            tc::tc_fence_after();
            {
                uint32_t d0[16], d1[16];
                tc::tmem_ld_32x16(tD + lane_sel + h2 * 16, d0);
                tc::tmem_ld_32x16(tD + lane_sel + 32 + h2 * 16, d1);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::bar_arrive_n(AR_BAR_EPI, 256 + 32);
                if (qrow < N1) {
                    const float inv = 1.0f / l;
                    const size_t ob = (size_t)(ao.seq(s) + (long long)qrow * ao.tok) * ao.ld + h * D + h2 * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        store_split4(Os, ob + i, make_float4((__uint_as_float(d0[i]) + __uint_as_float(d1[i])) * inv,
                                                             (__uint_as_float(d0[i + 1]) + __uint_as_float(d1[i + 1])) * inv,
                                                             (__uint_as_float(d0[i + 2]) + __uint_as_float(d1[i + 2])) * inv,
                                                             (__uint_as_float(d0[i + 3]) + __uint_as_float(d1[i + 3])) * inv));
                }
            }
        }
    }
    if (warp == 16) tc::bar_sync_n(AR_BAR_EPI, 256 + 32);   // pairs with the math threads' arrive after their LAST output read
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) tc::tmem_dealloc(tmem_base, 512);
}

// nseq sequences, H heads of 32; queries N1 (addr aq), keys/values N2 <= 448 (addr akv); split-bf16 output (addr ao).
static inline bool attn_rows_tc_supported(int N2) { return N2 >= 1 && N2 <= AR_MAXK; }

static inline int launch_attn_rows_tc(const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, SplitOut Os, AttnAddr ao, int nseq,
                                      int H, int N1, int N2, cudaStream_t st) {
    if (!pmce_configure_smem<attn_rows_tc_kernel>(AR_SMEM)) return 2;
    const long long work = (long long)nseq * H * ((N1 + 127) / 128);
    const int sms = tc_num_sms();
    const int grid = (int)(work < sms ? (work < 1 ? 1 : work) : sms);
    // Q carries softmax scale and log2e: the scores come out in the log2 domain
    attn_rows_tc_kernel<<<grid, AR_THREADS, AR_SMEM, st>>>(Q, aq, K, V, akv, Os, ao, N1, N2, nseq, H, 1.4426950408889634f / sqrtf(32.0f));
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
