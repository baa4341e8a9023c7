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
//   warp 8 (producer, one lane): the qkv projection already wrote q/k/v as split-bf16 [tokens, 3C] tensors, so the six
//       operand tiles of an item arrive by TMA (128-byte swizzle, straight into the layout UMMA expects): one 128-row box
//       per tensor in the spatial pass (tokens of a tile are contiguous), one T-row box per sequence in the temporal pass
//       (3-D view [frame][joint][3C]); completion is signalled on full[group].
// The block-diagonal trick spends 128/L x more MMA flops than the attention needs, which is free here: the CUDA-core
// kernel it replaces ran with 17 of 32 lanes active and ~3 warps per scheduler.
#pragma once
#include <stdlib.h>
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "gemm_tc.cuh"

constexpr int AT_TILE = 128 * 128;                 // bytes: [128 rows][64 bf16]
constexpr int AT_BUF = 6 * AT_TILE;                // Qh Ql Kh Kl Vh Vl   (P hi aliases Qh|Ql, P lo aliases Kh|Kl)
constexpr int AT_SMEM = 2 * AT_BUF + 1024 + 128;
constexpr int AT_THREADS = 288;                   // 8 math warps + 1 producer warp

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

namespace tc {
__device__ __forceinline__ void tma_store_3d_bulk(const CUtensorMap* m, uint32_t smem_src, int crd0, int crd1, int crd2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(crd0), "r"(crd1), "r"(crd2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d_bulk(const CUtensorMap* m, uint32_t smem_src, int crd0, int crd1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(crd0), "r"(crd1)
                 : "memory");
}
}  // namespace tc

// Output: each thread writes its O row (split-bf16) into the group's Q tiles (free once the PV MMAs are complete) in the
// swizzled tile layout and ONE thread stores the tile by TMA - spatial pass: one [G*L rows][64] box per half (the padding rows
// of the tile belong to the next tile and are not part of the box), temporal pass: one [T][1][64] box per sequence of the 3-D
// view. The smem buffer is handed back to the producer only after the store has read it. (Row-per-thread global stores cost
// 14 us of a 58 us spatial launch and 20 us of a 53 us temporal one: PMCE_ATTN_DEBUG=8.)
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tile_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_ohi,
                    const __grid_constant__ CUtensorMap tm_olo, int temporal, int C, int J, int T, int L, int G, int nseq, int H, float scale,
                    int dbg) {
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * AT_BUF);   // [2] per group, producer arrive.expect_tx + TMA bytes
    uint64_t* empty_bar = full_bar + 2;                                    // [2] per group, 1 arrival (tcgen05.commit)
    uint64_t* bar_s = empty_bar + 2;                                       // [2] S complete
    uint64_t* bar_o = bar_s + 2;                                           // [2] O complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_o + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&full_bar[b], 1); tc::mbar_init(&empty_bar[b], 1);
            tc::mbar_init(&bar_s[b], 1); tc::mbar_init(&bar_o[b], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 512);
    for (int i = tid; i < 2 * AT_BUF / 16; i += AT_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);   // stale tiles stay finite
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int ntiles = (nseq + G - 1) / G;
    pdl_trigger();
    pdl_wait();          // the prologue above overlapped the previous kernel; its outputs are visible from here
    const int nwork = ntiles * H;
    // this CTA's items: work = blockIdx.x + n * gridDim.x, n = 0, 1, ...; item n belongs to group n & 1

    if (warp >= 8) {
        // ================= producer: one lane issues the TMA loads =================
        if (tid == 256) {
            tc::tma_prefetch_desc(&tm_hi);
            tc::tma_prefetch_desc(&tm_lo);
            uint32_t n = 0;
            for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++n) {
                const int grp = n & 1;
                const uint32_t j = n >> 1;                                  // per-group item counter
                uint8_t* base_p = smem + grp * AT_BUF;
                const int tile = work / H, h = work - tile * H;
                tc::mbar_wait(&empty_bar[grp], (j & 1) ^ 1);
                if (dbg & 1) { tc::mbar_arrive(&full_bar[grp]); continue; }
                if (!temporal) {
                    // tokens of the tile are the 128 consecutive rows starting at tile*G*L (rows past G*L are masked, rows past the tensor are zero)
                    tc::mbar_arrive_expect_tx(&full_bar[grp], 6 * AT_TILE);
                    const int row0 = tile * G * L;
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        tc::tma_load_2d(base_p + (2 * x) * AT_TILE, &tm_hi, &full_bar[grp], x * C + h * 64, row0);
                        tc::tma_load_2d(base_p + (2 * x + 1) * AT_TILE, &tm_lo, &full_bar[grp], x * C + h * 64, row0);
                    }
                } else {
                    // sequence s = (b, j): its T frames are one [T][64] box of the 3-D view [b*T + t][j][3C], landing at rows [gl*T, gl*T+T)
                    const int s0 = tile * G;
                    const int ns = nseq - s0 < G ? nseq - s0 : G;
                    tc::mbar_arrive_expect_tx(&full_bar[grp], (uint32_t)(6 * ns * T * 128));
                    for (int gl = 0; gl < ns; ++gl) {
                        const int sq = s0 + gl, b = sq / J, jj = sq - b * J;
#pragma unroll
                        for (int x = 0; x < 3; ++x) {
                            tc::tma_load_3d(base_p + (2 * x) * AT_TILE + gl * T * 128, &tm_hi, &full_bar[grp], x * C + h * 64, jj, b * T);
                            tc::tma_load_3d(base_p + (2 * x + 1) * AT_TILE + gl * T * 128, &tm_lo, &full_bar[grp], x * C + h * 64, jj, b * T);
                        }
                    }
                }
            }
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
                        if (col >= cs && col < ce) m = fmaxf(m, __uint_as_float(v[i]) * scale);
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
                            p[i] = (col >= cs && col < ce) ? __expf(__uint_as_float(v[jj * 8 + i]) * scale - m) : 0.f;
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
            }
            tc::mbar_wait(&bar_o[grp], j & 1);
            tc::tc_fence_after();

            // ---- epilogue: O row / rowsum -> split bf16 -> staging tiles (the group's Q tiles) -> TMA store ----
            const float inv = 1.0f / lsum;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v0[32];
                tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v0);
                tc::tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = __uint_as_float(v0[jj * 8 + i]) * inv;
                    uint4 hh, ll;
                    tc::split8(o8, hh, ll);
                    tc::sts16(Qh, r, hf * 4 + jj, hh);
                    tc::sts16(Ql, r, hf * 4 + jj, ll);
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::bar_sync_group(1 + grp);
            if (issuer) {
                if (!(dbg & 8)) {
                    if (!temporal) {
                        tc::tma_store_2d_bulk(&tm_ohi, Qh, h * 64, tile * G * L);
                        tc::tma_store_2d_bulk(&tm_olo, Ql, h * 64, tile * G * L);
                    } else {
                        const int s0 = tile * G;
                        const int ns = nseq - s0 < G ? nseq - s0 : G;
                        for (int gl = 0; gl < ns; ++gl) {
                            const int sq = s0 + gl, b = sq / J, jj = sq - b * J;
                            tc::tma_store_3d_bulk(&tm_ohi, Qh + gl * T * 128, h * 64, jj, b * T);
                            tc::tma_store_3d_bulk(&tm_olo, Ql + gl * T * 128, h * 64, jj, b * T);
                        }
                    }
                    tc::tma_store_commit();
                    tc::tma_store_wait_read<0>();
                }
                tc::mbar_arrive(&empty_bar[grp]);                           // the producer may refill the buffer
            }
        }
        if (issuer) tc::tma_store_wait_read<0>();             // smem must outlive the reads; the grid boundary orders the writes
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// nseq sequences of L tokens (L <= 128; temporal: L % 8 == 0), H heads of 64. q|k|v are the split-bf16 [ntok, 3C] outputs of the
// qkv projection (token rows ordered (b, t, j)); spatial: sequence = (b,t), tokens j; temporal: sequence = (b,j), tokens t.
static inline int launch_attn_tile_tc(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, int ntok, int C, int J, int T, bool temporal, SplitOut Os,
                                      int nseq, int H, cudaStream_t st) {
    if (!pmce_configure_smem<attn_tile_tc_kernel>(AT_SMEM)) return 2;
    const int L = temporal ? T : J;
    const int G = 128 / L;
    CUtensorMap th, tl, toh, tol;      // Os: split-bf16 output [ntok, C], rows in the same (b, t, j) order as qkv
    if (!temporal) {
        if (make_tmap_bf16(&th, qkv_hi, ntok, 3 * C, 3 * C, 128) || make_tmap_bf16(&tl, qkv_lo, ntok, 3 * C, 3 * C, 128) ||
            make_tmap_bf16(&toh, Os.hi, ntok, C, C, G * L) || make_tmap_bf16(&tol, Os.lo, ntok, C, C, G * L))
            return 1;
    } else {
        if (make_tmap_bf16_3d(&th, qkv_hi, 3 * C, J, ntok / J, 3LL * C, 3LL * C * J, 1, T) ||
            make_tmap_bf16_3d(&tl, qkv_lo, 3 * C, J, ntok / J, 3LL * C, 3LL * C * J, 1, T) ||
            make_tmap_bf16_3d(&toh, Os.hi, C, J, ntok / J, (long long)C, (long long)C * J, 1, T) ||
            make_tmap_bf16_3d(&tol, Os.lo, C, J, ntok / J, (long long)C, (long long)C * J, 1, T))
            return 1;
    }
    const int ntiles = (nseq + G - 1) / G;
    const long long work = (long long)ntiles * H;
    const int sms = tc_num_sms();
    const long long want = (work + 1) / 2;                 // two items in flight per CTA
    const int grid = (int)(want < sms ? (want < 1 ? 1 : want) : sms);
    static int dbg = -1;   // PMCE_ATTN_DEBUG: profiling knobs (wrong results): 1 no loads, 2 no math, 4 no softmax, 8 no stores
    if (dbg < 0) dbg = pmce_profiling_knob("PMCE_ATTN_DEBUG");
    if (pmce_launch(attn_tile_tc_kernel, dim3(grid), dim3(AT_THREADS), AT_SMEM, st, 0, th, tl, toh, tol, temporal ? 1 : 0, C, J, T, L, G, nseq, H,
                    1.0f / sqrtf(64.0f), dbg) != cudaSuccess) return 3;
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
