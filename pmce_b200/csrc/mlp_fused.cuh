// Fused AdaLayerNorm + Mlp of the co-evolution decoder's 64-wide token streams (CoevoDecoder.py:86 / :104 with timm Mlp:
//     x <- x + fc2( GELU( fc1( AdaLN(x, g) ) ) ),   fc1: 64 -> 256, fc2: 256 -> 64 )
// one kernel instead of AdaLN-apply + fc1 GEMM (+GELU) + fc2 GEMM (+residual): the [rows, 256] hidden activations never leave
// the SM (they went out to HBM/L2 and back as split-bf16, 4 x the bytes of the stream itself), and the stream is read once and
// written once. Epilogue variants close the block without another launch:
//     MLP_EPI_X     x'' -> global (fp32)
//     MLP_EPI_F2C   coords_out = W_f2c x'' + b + coords_in   (CoevoDecoder.py:189) - x'' itself is never written
//     MLP_EPI_T     x'' -> global and t = AdaLN_next(x'') as split-bf16 (the A operand of the next projection, e.g. qkv)
//
// Work item = 128 consecutive rows of the flat [rows, 64] stream (rows of different clips pick their own gamma/beta).
// fc1 and fc2 weights (split-bf16, 128 KB) are loaded ONCE per persistent CTA and stay in shared memory.
//   warps 0-3   row owners: thread r keeps the whole row in registers (thread-local LayerNorm statistics, residual), writes the
//               A tiles, and runs half of the GELU work
//   warps 4-7   GELU helpers: same TMEM lanes as warp w-4, the other 32 columns of each 64-wide hidden chunk
//   warp 8      TMA producer (weights once, then the x tile of the next item while the current one is computed)
//   warp 9      MMA issuer: fc1 as four N=64 chunks (each with its own completion barrier, so GELU on chunk 0 starts while
//               chunks 1-3 are still in the tensor core), fc2 as four K=64 partial products accumulated in TMEM as the hidden
//               chunks arrive (two hidden buffers: GELU of chunk c+1 overlaps the fc2 MMAs of chunk c)
// Hand-offs between the compute warps and the issuer are named barriers (bar.arrive / bar.sync), completions are mbarriers
// signalled by tcgen05.commit. TMEM: 4 x 64 columns hidden + 64 columns output.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"
#include "ca_fused.cuh"

enum MlpEpi { MLP_EPI_X = 0, MLP_EPI_F2C = 1, MLP_EPI_T = 2 };

constexpr int MLP_THREADS = 320;                    // 8 compute warps + producer + MMA issuer
constexpr int MLP_W1 = 2 * 256 * 128;               // fc1 weight [256][64] bf16, hi | lo
constexpr int MLP_W2 = 2 * 4 * 64 * 128;            // fc2 weight as 4 K-blocks of [64][64] bf16, hi (4 blocks) | lo (4 blocks)
constexpr int MLP_OFF_W2 = MLP_W1;
constexpr int MLP_OFF_IN = MLP_W1 + MLP_W2;         // x tile: two [128][32 fp32] boxes
constexpr int MLP_OFF_A = MLP_OFF_IN + 32768;       // A tiles hi | lo; hidden buffer 1 after fc1; t tiles for the store (EPI_T)
constexpr int MLP_OFF_H = MLP_OFF_A + 32768;        // hidden buffer 0 (hi | lo); fp32 output boxes for the store
constexpr int MLP_OFF_BAR = MLP_OFF_H + 32768;
constexpr int MLP_SMEM = MLP_OFF_BAR + 256 + 1024;
static_assert(MLP_SMEM <= 227 * 1024, "shared memory");

namespace tc {
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
}  // namespace tc

struct MlpFusedArgs {
    float* x;                 // [rows, 64] fp32, updated in place (not written for MLP_EPI_F2C)
    const float* gb;          // [B, gb_ld] AdaLN gamma|beta of every slot
    int gb_ld, slot, slot_next;
    const float *b1, *b2;     // [256], [64]
    int rows, ntok;           // rows = B * ntok
    float eps;
    // MLP_EPI_F2C
    const float *wc, *bc;     // [3,64], [3]
    const float* coords_in;   // [rows, 3]
    float* coords_out;        // [rows, 3]
};

// named barrier ids (0 = __syncthreads)
constexpr int MLP_BAR_OWN = 1;      // the 128 row owners
constexpr int MLP_BAR_T = 2;        // owners arrive, issuer syncs: A tiles written
constexpr int MLP_BAR_H0 = 3;       // compute warps arrive, issuer syncs: hidden buffer 0 / 1 written (ids 3, 4)

// AdaLayerNorm of a row held in registers -> split-bf16 A tiles (CoevoDecoder.py:23-29: unbiased std, eps added to the std)
__device__ __forceinline__ void mlp_adaln_to_tiles(const float (&x)[64], const float* __restrict__ g, float eps, uint32_t t_hi, uint32_t t_lo, int r) {
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
    const float inv = 1.0f / (sqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / 63.0f)) + eps);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + 8 * c)), g1 = __ldg(reinterpret_cast<const float4*>(g + 8 * c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(g + 64 + 8 * c)), b1 = __ldg(reinterpret_cast<const float4*>(g + 64 + 8 * c + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = gg[i] * (x[8 * c + i] - mean) * inv + bb[i];
        uint4 hh, ll;
        tc::split8(y, hh, ll);
        tc::sts16(t_hi, r, c, hh);
        tc::sts16(t_lo, r, c, ll);
    }
}

template <int EPI>
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp64_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                   const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                   const __grid_constant__ CUtensorMap tm_t_hi, const __grid_constant__ CUtensorMap tm_t_lo, MlpFusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MLP_OFF_BAR);
    uint64_t* w_full = bars;            // weights landed (once)
    uint64_t* x_full = bars + 1;        // x tile landed
    uint64_t* x_empty = bars + 2;       // 4 arrivals: the owner warps hold their rows in registers
    uint64_t* fc1_done = bars + 3;      // [4] tcgen05.commit per hidden chunk
    uint64_t* h_free = bars + 7;        // [2] tcgen05.commit: the fc2 MMAs of chunk 0 / 1 have read hidden buffer 0 / 1
    uint64_t* fc2_done = bars + 9;      // tcgen05.commit: output accumulator complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (a.rows + 127) / 128;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tm_x); tc::tma_prefetch_desc(&tm_w1_hi); tc::tma_prefetch_desc(&tm_w1_lo);
        tc::tma_prefetch_desc(&tm_w2_hi); tc::tma_prefetch_desc(&tm_w2_lo);
        if (EPI == MLP_EPI_T) { tc::tma_prefetch_desc(&tm_t_hi); tc::tma_prefetch_desc(&tm_t_lo); }
        tc::mbar_init(w_full, 1); tc::mbar_init(x_full, 1); tc::mbar_init(x_empty, 4);
        for (int c = 0; c < 4; ++c) tc::mbar_init(&fc1_done[c], 1);
        tc::mbar_init(&h_free[0], 1); tc::mbar_init(&h_free[1], 1); tc::mbar_init(fc2_done, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 8) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tH = tmem_base, tO = tmem_base + 256;
    pdl_trigger();
    if (warp != 8) pdl_wait();     // the producer first requests the weights (constants), which then land under the previous kernel's tail

    if (warp == 8) {
        // ================= producer =================
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, MLP_W1 + MLP_W2);
            tc::tma_load_2d(smem, &tm_w1_hi, w_full, 0, 0);
            tc::tma_load_2d(smem + MLP_W1 / 2, &tm_w1_lo, w_full, 0, 0);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                tc::tma_load_2d(smem + MLP_OFF_W2 + kb * 8192, &tm_w2_hi, w_full, kb * 64, 0);
                tc::tma_load_2d(smem + MLP_OFF_W2 + MLP_W2 / 2 + kb * 8192, &tm_w2_lo, w_full, kb * 64, 0);
            }
            pdl_wait();            // x is the previous kernel's output
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                tc::mbar_wait(x_empty, (it & 1) ^ 1);
                tc::mbar_arrive_expect_tx(x_full, 32768);
                tc::tma_load_2d(smem + MLP_OFF_IN, &tm_x, x_full, 0, tile * 128);
                tc::tma_load_2d(smem + MLP_OFF_IN + 16384, &tm_x, x_full, 32, tile * 128);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer (whole warp walks the barriers, one lane issues) =================
        uint32_t it = 0;
        tc::mbar_wait(w_full, 0);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            tc::bar_sync_n(MLP_BAR_T, 128 + 32);                       // A tiles written (and the previous item's TMEM reads are done)
            if (lane == 0) {
                tc::tc_fence_after();
                constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 64);
                const uint64_t ah = tc::umma_desc_sw128(sb + MLP_OFF_A), al = tc::umma_desc_sw128(sb + MLP_OFF_A + AT_TILE);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint64_t wh = tc::umma_desc_sw128(sb + c * 8192), wl = tc::umma_desc_sw128(sb + MLP_W1 / 2 + c * 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tH + c * 64, tc::umma_desc_advance_k(al, k), tc::umma_desc_advance_k(wh, k), idesc, k != 0);
                        tc::umma_bf16(tH + c * 64, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                        tc::umma_bf16(tH + c * 64, tc::umma_desc_advance_k(ah, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
                    }
                    tc::umma_commit(&fc1_done[c]);
                }
            }
            __syncwarp();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tc::bar_sync_n(MLP_BAR_H0 + (c & 1), 256 + 32);        // hidden chunk c written to buffer c & 1
                if (lane == 0) {
                    tc::tc_fence_after();
                    constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 64);
                    const uint32_t hb = sb + ((c & 1) ? MLP_OFF_A : MLP_OFF_H);
                    const uint64_t hh = tc::umma_desc_sw128(hb), hl = tc::umma_desc_sw128(hb + AT_TILE);
                    const uint64_t wh = tc::umma_desc_sw128(sb + MLP_OFF_W2 + c * 8192), wl = tc::umma_desc_sw128(sb + MLP_OFF_W2 + MLP_W2 / 2 + c * 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tO, tc::umma_desc_advance_k(hl, k), tc::umma_desc_advance_k(wh, k), idesc, (c | k) != 0);
                        tc::umma_bf16(tO, tc::umma_desc_advance_k(hh, k), tc::umma_desc_advance_k(wl, k), idesc, 1);
                        tc::umma_bf16(tO, tc::umma_desc_advance_k(hh, k), tc::umma_desc_advance_k(wh, k), idesc, 1);
                    }
                    if (c < 2) tc::umma_commit(&h_free[c]);        // hidden buffer c is rewritten by chunk c + 2 (the only waiter)
                    if (c == 3) tc::umma_commit(fc2_done);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= compute warps =================
        const bool owner = warp < 4;
        const int r = tid & 127, half = warp >> 2;
        const int sw = r & 7;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t in_row = sb + MLP_OFF_IN + r * 128, out_row = sb + MLP_OFF_H + r * 128;
        bool store_pending = false;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int row = tile * 128 + r;
            float x[64];
            if (owner) {
                // ---- the row -> registers; the IN buffer goes back to the producer (next item's rows stream in meanwhile) ----
                tc::mbar_wait(x_full, it & 1);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float4 v = tc::lds16(in_row + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4));
                    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(x_empty);
                // the A / H buffers are free once the previous item's TMA stores have read them
                if (tid == 0 && store_pending) tc::tma_store_wait_read<0>();
                tc::bar_sync_n(MLP_BAR_OWN, 128);
                const int rr = row < a.rows ? row : a.rows - 1;                       // rows past the end: any valid gamma/beta (the rows are zero)
                const float* g = a.gb + (size_t)(rr / a.ntok) * a.gb_ld + a.slot * 128;
                mlp_adaln_to_tiles(x, g, a.eps, sb + MLP_OFF_A, sb + MLP_OFF_A + AT_TILE, r);
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::bar_arrive_n(MLP_BAR_T, 128 + 32);
            }
            // ---- GELU on the four hidden chunks: this thread's 32 of the chunk's 64 columns ----
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tc::mbar_wait(&fc1_done[c], it & 1);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld_32x32(tH + lane_sel + c * 64 + half * 32, v);
                tc::tmem_ld_wait();
                // buffer c & 1 must be free: buffer 0 (H) after the fc2 MMAs of chunk c-2; buffer 1 (A) after ALL fc1 MMAs (they
                // read the A tiles) and, from chunk 3 on, after the fc2 MMAs of chunk 1
                if (c == 1) tc::mbar_wait(&fc1_done[3], it & 1);
                if (c >= 2) tc::mbar_wait(&h_free[c & 1], it & 1);     // the fc2 MMAs of chunk c - 2 have read the buffer (one commit per item)
                const uint32_t hb = sb + ((c & 1) ? MLP_OFF_A : MLP_OFF_H);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(a.b1 + c * 64 + half * 32 + cc * 8));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b1 + c * 64 + half * 32 + cc * 8 + 4));
                    float y[8];
                    y[0] = gelu_erf(__uint_as_float(v[cc * 8]) + ba.x); y[1] = gelu_erf(__uint_as_float(v[cc * 8 + 1]) + ba.y);
                    y[2] = gelu_erf(__uint_as_float(v[cc * 8 + 2]) + ba.z); y[3] = gelu_erf(__uint_as_float(v[cc * 8 + 3]) + ba.w);
                    y[4] = gelu_erf(__uint_as_float(v[cc * 8 + 4]) + bb.x); y[5] = gelu_erf(__uint_as_float(v[cc * 8 + 5]) + bb.y);
                    y[6] = gelu_erf(__uint_as_float(v[cc * 8 + 6]) + bb.z); y[7] = gelu_erf(__uint_as_float(v[cc * 8 + 7]) + bb.w);
                    uint4 hh, ll;
                    tc::split8(y, hh, ll);
                    tc::sts16(hb, r, half * 4 + cc, hh);
                    tc::sts16(hb + AT_TILE, r, half * 4 + cc, ll);
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::bar_arrive_n(MLP_BAR_H0 + (c & 1), 256 + 32);
            }
            if (owner) {
                // ---- x'' = x + fc2(...) + b2 ----
                tc::mbar_wait(fc2_done, it & 1);
                tc::tc_fence_after();
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tO + lane_sel + hf * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.b2 + hf * 32 + i));
                        x[hf * 32 + i] += __uint_as_float(v[i]) + b2.x; x[hf * 32 + i + 1] += __uint_as_float(v[i + 1]) + b2.y;
                        x[hf * 32 + i + 2] += __uint_as_float(v[i + 2]) + b2.z; x[hf * 32 + i + 3] += __uint_as_float(v[i + 3]) + b2.w;
                    }
                }
                if (EPI == MLP_EPI_F2C) {
                    // coords_out = W_f2c x'' + b + coords_in (CoevoDecoder.py:189): three 64-long dot products per row, thread-local
                    if (row < a.rows) {
                        float o3[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll
                            for (int i = 0; i < 64; i += 4) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(a.wc + k * 64 + i));
                                a0 = fmaf(x[i], w.x, a0); a1 = fmaf(x[i + 1], w.y, a1); a0 = fmaf(x[i + 2], w.z, a0); a1 = fmaf(x[i + 3], w.w, a1);
                            }
                            o3[k] = (a0 + a1) + __ldg(a.bc + k) + a.coords_in[(size_t)row * 3 + k];
                        }
                        a.coords_out[(size_t)row * 3] = o3[0]; a.coords_out[(size_t)row * 3 + 1] = o3[1]; a.coords_out[(size_t)row * 3 + 2] = o3[2];
                    }
                    tc::tc_fence_before();                                   // (the issuer's next MLP_BAR_T sync orders the TMEM reads above)
                } else {
                    // the hidden buffers are free (fc2_done covers every MMA): x'' -> fp32 boxes in H, t -> tiles in A
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        tc::sts16f(out_row + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
                    if (EPI == MLP_EPI_T) {
                        const int rr = row < a.rows ? row : a.rows - 1;
                        const float* g = a.gb + (size_t)(rr / a.ntok) * a.gb_ld + a.slot_next * 128;
                        mlp_adaln_to_tiles(x, g, a.eps, sb + MLP_OFF_A, sb + MLP_OFF_A + AT_TILE, r);
                    }
                    tc::fence_proxy_async();
                    tc::tc_fence_before();
                    tc::bar_sync_n(MLP_BAR_OWN, 128);
                    if (tid == 0) {
                        tc::tma_store_2d(&tm_x, smem + MLP_OFF_H, 0, tile * 128);
                        tc::tma_store_2d(&tm_x, smem + MLP_OFF_H + 16384, 32, tile * 128);
                        if (EPI == MLP_EPI_T) {
                            tc::tma_store_2d(&tm_t_hi, smem + MLP_OFF_A, 0, tile * 128);
                            tc::tma_store_2d(&tm_t_lo, smem + MLP_OFF_A + AT_TILE, 0, tile * 128);
                        }
                        tc::tma_store_commit();
                        store_pending = true;
                    }
                }
            }
        }
        if (tid == 0 && store_pending) tc::tma_store_wait_read<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_base, 512);
}

// ---- host side -------------------------------------------------------------------------------------------
struct MlpWeights {   // split-bf16 fc1 [256,64] and fc2 [64,256] (row-major, K contiguous)
    const __nv_bfloat16 *w1_hi, *w1_lo, *w2_hi, *w2_lo;
};

template <int EPI>
static inline int launch_mlp64_fused_t(const CUtensorMap* m, const MlpFusedArgs& a, cudaStream_t st) {
    if (!pmce_configure_smem<mlp64_fused_kernel<EPI>>(MLP_SMEM)) return 2;
    const int ntiles = (a.rows + 127) / 128;
    const int cap = tc_num_sms();
    return pmce_launch(mlp64_fused_kernel<EPI>, dim3(ntiles < cap ? ntiles : cap), dim3(MLP_THREADS), MLP_SMEM, st, 0,
                       m[0], m[1], m[2], m[3], m[4], m[5], m[6], a) == cudaSuccess ? 0 : 3;
}

// x [rows, 64] fp32 in place; t (MLP_EPI_T only): split-bf16 [rows, 64]
static inline int launch_mlp64_fused(int epi, const MlpWeights& w, SplitOut t, const MlpFusedArgs& a, cudaStream_t st) {
    CUtensorMap m[7];
    if (make_tmap(&m[0], a.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.rows, 64, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap_bf16(&m[1], w.w1_hi, 256, 64, 64, 256) || make_tmap_bf16(&m[2], w.w1_lo, 256, 64, 64, 256) ||
        make_tmap_bf16(&m[3], w.w2_hi, 64, 256, 256, 64) || make_tmap_bf16(&m[4], w.w2_lo, 64, 256, 256, 64))
        return 1;
    if (epi == MLP_EPI_T) {
        if (make_tmap_bf16(&m[5], t.hi, a.rows, 64, 64, 128) || make_tmap_bf16(&m[6], t.lo, a.rows, 64, 64, 128)) return 1;
    } else {
        m[5] = m[1]; m[6] = m[2];
    }
    if (epi == MLP_EPI_X) return launch_mlp64_fused_t<MLP_EPI_X>(m, a, st);
    if (epi == MLP_EPI_F2C) return launch_mlp64_fused_t<MLP_EPI_F2C>(m, a, st);
    return launch_mlp64_fused_t<MLP_EPI_T>(m, a, st);
}
