// One GRU time step on tcgen05: gh = h_prev W_hh^T for U hidden units x 3 gates per CTA (N = 3U accumulator columns:
// [0,U) r, [U,2U) z, [2U,3U) n), batch rows on the 128 TMEM lanes, K = H streamed by TMA; the gate math
//   r = s(gi_r + gh_r + b_r), z = s(gi_z + gh_z + b_z), n = tanh(gi_n + r (gh_n + b_n)), h' = (1-z) n + z h
// (PyTorch nn.GRU convention, rows of W_hh ordered r,z,n) is fused in the TMEM epilogue, which also emits the
// split-bf16 copy of h' that the next step's MMA (and the layer-1 input projection) consumes.
// W_hh is used in place: the three gate row-blocks {j0, H+j0, 2H+j0} are three TMA boxes stacked in shared memory.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"

struct GruTcDir {
    const float* gi;      // [B, >=3H] input projections for this step (ld_gi)
    const float* hprev;   // fp32 [B, H] (ld_h) for the z*h term
    const float* bhh;     // [3H]
    float* hout;          // fp32 [B, H] (ld_o)
    SplitOut hs;          // split copy of h' (ld_s) or nulls
    int ld_gi, ld_h, ld_o, ld_s;
};


constexpr int GRU_U = 16;                       // hidden units per CTA
constexpr int GRU_N = 3 * GRU_U;                // accumulator columns
constexpr int GRU_A_TILE = 128 * 128;           // bytes (128 rows x 64 bf16)
constexpr int GRU_W_TILE = GRU_N * 128;
constexpr int GRU_STAGE = 2 * GRU_A_TILE + 2 * GRU_W_TILE;
constexpr int GRU_STAGES = 4;
constexpr int GRU_SMEM = GRU_STAGES * GRU_STAGE + 1024 + 256;

struct GruTcMaps {   // per direction: h_prev (hi, lo) [B, H] and W_hh (hi, lo) [3H, H]
    CUtensorMap h_hi, h_lo, w_hi, w_lo;
};

__global__ void __launch_bounds__(192, 1)
gru_step_tc_kernel(const __grid_constant__ GruTcMaps m0_, const __grid_constant__ GruTcMaps m1_, GruTcDir d0, GruTcDir d1, int B, int H) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GRU_STAGES * GRU_STAGE);
    uint64_t* empty_bar = full_bar + GRU_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GRU_STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const GruTcMaps& mp = dir == 0 ? m0_ : m1_;
    const GruTcDir d = dir == 0 ? d0 : d1;
    const int j0 = blockIdx.x * GRU_U, b0 = blockIdx.z * 128;
    const int nkb = H / 64;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mp.h_hi); tc::tma_prefetch_desc(&mp.h_lo);
        tc::tma_prefetch_desc(&mp.w_hi); tc::tma_prefetch_desc(&mp.w_lo);
        for (int s = 0; s < GRU_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr_smem, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();
    if (threadIdx.x != 0) pdl_wait();   // the producer first requests the W_hh tiles of the ring's first pass (constants)

    if (warp == 0) {
        if (lane == 0) {
            auto load_w = [&](int kb, int s) {
                uint8_t* wh = smem + s * GRU_STAGE + 2 * GRU_A_TILE;
                uint8_t* wl = wh + GRU_W_TILE;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    tc::tma_load_2d(wh + g * GRU_U * 128, &mp.w_hi, &full_bar[s], kb * 64, g * H + j0);
                    tc::tma_load_2d(wl + g * GRU_U * 128, &mp.w_lo, &full_bar[s], kb * 64, g * H + j0);
                }
            };
            auto load_h = [&](int kb, int s) {
                uint8_t* st = smem + s * GRU_STAGE;
                tc::tma_load_2d(st, &mp.h_hi, &full_bar[s], kb * 64, b0);
                tc::tma_load_2d(st + GRU_A_TILE, &mp.h_lo, &full_bar[s], kb * 64, b0);
            };
            // first pass over the ring (every stage is free): the weight tiles go out before the previous step has finished,
            // h_prev (its output) after
            const int pre = nkb < GRU_STAGES ? nkb : GRU_STAGES;
            for (int kb = 0; kb < pre; ++kb) { tc::mbar_arrive_expect_tx(&full_bar[kb], GRU_STAGE); load_w(kb, kb); }
            pdl_wait();
            for (int kb = 0; kb < pre; ++kb) load_h(kb, kb);
            for (int kb = pre; kb < nkb; ++kb) {
                const int s = kb % GRU_STAGES;
                const uint32_t ph = (kb / GRU_STAGES) & 1;
                tc::mbar_wait(&empty_bar[s], ph ^ 1);
                tc::mbar_arrive_expect_tx(&full_bar[s], GRU_STAGE);
                load_h(kb, s);
                load_w(kb, s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, GRU_N);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % GRU_STAGES;
                const uint32_t ph = (kb / GRU_STAGES) & 1;
                tc::mbar_wait(&full_bar[s], ph);
                tc::tc_fence_after();
                const uint32_t st = tc::smem_u32(smem + s * GRU_STAGE);
                const uint64_t a_hi = tc::umma_desc_sw128(st), a_lo = tc::umma_desc_sw128(st + GRU_A_TILE);
                const uint64_t w_hi = tc::umma_desc_sw128(st + 2 * GRU_A_TILE), w_lo = tc::umma_desc_sw128(st + 2 * GRU_A_TILE + GRU_W_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                    tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                    tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                }
                tc::umma_commit(&empty_bar[s]);
            }
            tc::umma_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int row = b0 + q * 32 + lane;
        tc::mbar_wait(tmem_full_bar, 0);
        tc::tc_fence_after();
        uint32_t ar[16], az[16], an[16];
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
        tc::tmem_ld_32x16(t0, ar);
        tc::tmem_ld_32x16(t0 + GRU_U, az);
        tc::tmem_ld_32x16(t0 + 2 * GRU_U, an);
        tc::tmem_ld_wait();
        if (row < B) {
            const float* gi = d.gi + (size_t)row * d.ld_gi + j0;
            const float* hp = d.hprev + (size_t)row * d.ld_h + j0;
            float hn[16];
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
                const float4 gr = ld4(gi + u), gz = ld4(gi + H + u), gn = ld4(gi + 2 * H + u), hv = ld4(hp + u);
                const float4 br = ld4(d.bhh + j0 + u), bz = ld4(d.bhh + H + j0 + u), bn = ld4(d.bhh + 2 * H + j0 + u);
                const float grr[4] = {gr.x, gr.y, gr.z, gr.w}, gzz[4] = {gz.x, gz.y, gz.z, gz.w}, gnn[4] = {gn.x, gn.y, gn.z, gn.w};
                const float hh[4] = {hv.x, hv.y, hv.z, hv.w}, brr[4] = {br.x, br.y, br.z, br.w}, bzz[4] = {bz.x, bz.y, bz.z, bz.w}, bnn[4] = {bn.x, bn.y, bn.z, bn.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float r = sigmoid_f(grr[i] + (__uint_as_float(ar[u + i]) + brr[i]));
                    const float z = sigmoid_f(gzz[i] + (__uint_as_float(az[u + i]) + bzz[i]));
                    const float n = tanhf(gnn[i] + r * (__uint_as_float(an[u + i]) + bnn[i]));
                    hn[u + i] = (1.0f - z) * n + z * hh[i];
                }
            }
            float* ho = d.hout + (size_t)row * d.ld_o + j0;
#pragma unroll
            for (int u = 0; u < 16; u += 4) st4(ho + u, make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]));
            if (d.hs.hi) {
                const size_t si = (size_t)row * d.ld_s + j0;
#pragma unroll
                for (int u = 0; u < 16; u += 4) store_split4(d.hs, si + u, make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------
// The same step on a FEW SMs (opt-in experiment, api.cu::gru_few_plan has the measurements): the kernel above spreads a step over
// H/16 x ndir = 128 CTAs for ~17 us. This variant runs it on `gridDim.x` CTAs (launched as CTA pairs so they fill whole TPCs):
// every CTA walks tiles of U hidden units x 3 gates (N = 3U) of one direction, two TMEM accumulators so the gate math of tile i
// overlaps the MMAs of tile i+1. On 12 CTAs a step takes ~60-70 us (per-SM L2 ingest) and, beside the lifter, costs the forward
// as much as the 128-CTA step - both are L2-bandwidth, not SM-count, problems.
// ------------------------------------------------------------------------------------------------------
template <int U, int S>
struct GruMultiCfg {
    static constexpr int N = 3 * U;
    static constexpr int W_TILE = N * 128;
    static constexpr int STAGE = 2 * GRU_A_TILE + 2 * W_TILE;
    static constexpr int SMEM = S * STAGE + 1024 + 256;
    static constexpr int ACC_STRIDE = N <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
};

template <int U, int S>
__global__ void __launch_bounds__(192, 1)
gru_step_multi_kernel(const __grid_constant__ GruTcMaps m0_, const __grid_constant__ GruTcMaps m1_, GruTcDir d0, GruTcDir d1, int B, int H, int ndir) {
    using Cfg = GruMultiCfg<U, S>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* tmem_full_bar = empty_bar + S;         // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = H / 64, jtiles = H / U;
    const int ntiles = jtiles * ndir * ((B + 127) / 128);     // tile = (batch block, direction, unit tile), unit tile fastest

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&m0_.h_hi); tc::tma_prefetch_desc(&m0_.h_lo); tc::tma_prefetch_desc(&m0_.w_hi); tc::tma_prefetch_desc(&m0_.w_lo);
        if (ndir > 1) { tc::tma_prefetch_desc(&m1_.h_hi); tc::tma_prefetch_desc(&m1_.h_lo); tc::tma_prefetch_desc(&m1_.w_hi); tc::tma_prefetch_desc(&m1_.w_lo); }
        for (int s = 0; s < S; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], 4); }
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int jt = tile % jtiles, dir = (tile / jtiles) % ndir, b0 = (tile / (jtiles * ndir)) * 128;
                const GruTcMaps& mp = dir == 0 ? m0_ : m1_;
                const int j0 = jt * U;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S;
                    tc::mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
                    uint8_t* st = smem + s * Cfg::STAGE;
                    tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE);
                    tc::tma_load_2d(st, &mp.h_hi, &full_bar[s], kb * 64, b0);
                    tc::tma_load_2d(st + GRU_A_TILE, &mp.h_lo, &full_bar[s], kb * 64, b0);
                    uint8_t* wh = st + 2 * GRU_A_TILE;
                    uint8_t* wl = wh + Cfg::W_TILE;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        tc::tma_load_2d(wh + g * U * 128, &mp.w_hi, &full_bar[s], kb * 64, g * H + j0);
                        tc::tma_load_2d(wl + g * U * 128, &mp.w_lo, &full_bar[s], kb * 64, g * H + j0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, Cfg::N);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                tc::mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);     // the gate math of tile tcount-2 has read this accumulator
                tc::tc_fence_after();
                const uint32_t dt = tmem_base + acc * Cfg::ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S;
                    tc::mbar_wait(&full_bar[s], (it / S) & 1);
                    tc::tc_fence_after();
                    const uint32_t st = tc::smem_u32(smem + s * Cfg::STAGE);
                    const uint64_t a_hi = tc::umma_desc_sw128(st), a_lo = tc::umma_desc_sw128(st + GRU_A_TILE);
                    const uint64_t w_hi = tc::umma_desc_sw128(st + 2 * GRU_A_TILE), w_lo = tc::umma_desc_sw128(st + 2 * GRU_A_TILE + Cfg::W_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(dt, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                        tc::umma_bf16(dt, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                        tc::umma_bf16(dt, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                    }
                    tc::umma_commit(&empty_bar[s]);
                }
                tc::umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const int jt = tile % jtiles, dir = (tile / jtiles) % ndir, b0 = (tile / (jtiles * ndir)) * 128;
            const GruTcDir& d = dir == 0 ? d0 : d1;
            const uint32_t acc = tcount & 1;
            const int row = b0 + q * 32 + lane;
            tc::mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t t0 = tmem_base + acc * Cfg::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
            if (b0 + q * 32 < B) {          // warps whose 32 rows are all padding only hand the accumulator back
#pragma unroll 1
                for (int u0 = 0; u0 < U; u0 += 16) {
                    uint32_t ar[16], az[16], an[16];
                    tc::tmem_ld_32x16(t0 + u0, ar);
                    tc::tmem_ld_32x16(t0 + U + u0, az);
                    tc::tmem_ld_32x16(t0 + 2 * U + u0, an);
                    tc::tmem_ld_wait();
                    if (row < B) {
                        const int j0 = jt * U + u0;
                        const float* gi = d.gi + (size_t)row * d.ld_gi + j0;
                        const float* hp = d.hprev + (size_t)row * d.ld_h + j0;
                        float hn[16];
#pragma unroll
                        for (int u = 0; u < 16; u += 4) {
                            const float4 gr = ld4(gi + u), gz = ld4(gi + H + u), gn = ld4(gi + 2 * H + u), hv = ld4(hp + u);
                            const float4 br = ld4(d.bhh + j0 + u), bz = ld4(d.bhh + H + j0 + u), bn = ld4(d.bhh + 2 * H + j0 + u);
                            const float grr[4] = {gr.x, gr.y, gr.z, gr.w}, gzz[4] = {gz.x, gz.y, gz.z, gz.w}, gnn[4] = {gn.x, gn.y, gn.z, gn.w};
                            const float hh[4] = {hv.x, hv.y, hv.z, hv.w}, brr[4] = {br.x, br.y, br.z, br.w}, bzz[4] = {bz.x, bz.y, bz.z, bz.w}, bnn[4] = {bn.x, bn.y, bn.z, bn.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float r = sigmoid_f(grr[i] + (__uint_as_float(ar[u + i]) + brr[i]));
                                const float z = sigmoid_f(gzz[i] + (__uint_as_float(az[u + i]) + bzz[i]));
                                const float n = tanhf(gnn[i] + r * (__uint_as_float(an[u + i]) + bnn[i]));
                                hn[u + i] = (1.0f - z) * n + z * hh[i];
                            }
                        }
                        float* ho = d.hout + (size_t)row * d.ld_o + j0;
#pragma unroll
                        for (int u = 0; u < 16; u += 4) st4(ho + u, make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]));
                        if (d.hs.hi) {
                            const size_t si = (size_t)row * d.ld_s + j0;
#pragma unroll
                            for (int u = 0; u < 16; u += 4) store_split4(d.hs, si + u, make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]));
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------
// One GRU LAYER (up to two directions) as ONE persistent launch: the per-step kernel above paid a launch, a pipeline fill and a
// TMEM allocation per time step (41 launches per forward, 17 us each for ~6 us of L2 streaming). Here the CTAs stay resident
// over all steps of the layer: CTA (x, dir) owns hidden units [16 x, 16 x + 16) of direction dir for every step,
//   warp 0   TMA producer: per step streams the 16 k-blocks of h_prev (all hidden units of the previous step, written by the
//            OTHER CTAs of this direction) and of its own W_hh rows (L2 hits); W_hh tiles of the first stages are requested
//            BEFORE the step barrier, the h_prev tiles after it
//   warp 1   MMA issuer, one TMEM accumulator (128 x 48), handed back by the epilogue through an mbarrier
//   warps 2-5 epilogue: gate math in registers (h_prev of the CTA's own units is its own output of the previous step), h' as
//            fp32 + split-bf16 into the layer's hidden-state tensor, then the step barrier: every CTA owns ONE flag word and
//            publishes "step s done" with st.release.gpu (no atomics, no __threadfence: a first version with one atomic counter
//            per direction spent 12 us per step in the fence behind 64 CTAs polling the same L2 line); the producer warp of every
//            CTA polls the direction's 64 flags (two ld.acquire.gpu per lane, __all_sync), then fence.proxy.async before its TMA
//            reads of h'.
// The two directions of a layer are independent chains with their own flags. All CTAs of a launch must be co-resident
// (64 per direction, 1 per SM: 128 <= 148 SMs); the host launches one batch tile of 128 rows at a time.
// Hidden states live in [nblk][B][ld] tensors (a block per time step) so the TMA boxes zero-fill the rows past B.
// ------------------------------------------------------------------------------------------------------
struct GruLayerDir {
    const float* gi;          // input projections: row b of step s at gi + s * gi_step + b * ld_gi (+ gate * H + unit)
    long long gi_step;
    int ld_gi;
    float* y;                 // hidden states fp32 [nblk, B, ld_y]; this direction's units start at column col0
    SplitOut ys;              // the same, split-bf16 (the A operand of the next step, read by TMA)
    int ld_y, col0;
    int blk0, blk_step;       // step s writes block blk0 + s * blk_step and reads block blk0 + (s - 1) * blk_step
    int nsteps;
    const float* bhh;         // [3H]
    float* last_out;          // optional [B, ld_last]: the last step's h' is also written here (columns = unit index)
    int ld_last;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr int GRU_MAX_CTAS = 64;     // CTAs (flags) per direction: H / 16 <= 64

__global__ void __launch_bounds__(192, 1)
gru_layer_tc_kernel(const __grid_constant__ GruTcMaps m0_, const __grid_constant__ GruTcMaps m1_, GruLayerDir d0, GruLayerDir d1, int B, int H, int b0,
                    unsigned* __restrict__ counters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GRU_STAGES * GRU_STAGE);
    uint64_t* empty_bar = full_bar + GRU_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GRU_STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const GruTcMaps& mp = dir == 0 ? m0_ : m1_;
    const GruLayerDir d = dir == 0 ? d0 : d1;
    const int j0 = blockIdx.x * GRU_U;
    const int nkb = H / 64;
    unsigned* flags = counters + dir * GRU_MAX_CTAS;     // flags[x] = number of steps CTA x of this direction has published
    const int nx = gridDim.x;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mp.h_hi); tc::tma_prefetch_desc(&mp.h_lo);
        tc::tma_prefetch_desc(&mp.w_hi); tc::tma_prefetch_desc(&mp.w_lo);
        for (int s = 0; s < GRU_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(tmem_full_bar, 1); tc::mbar_init(tmem_empty_bar, 4);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr_smem, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        uint32_t it = 0;
        for (int s = 1; s < d.nsteps; ++s) {
            const int blkp = d.blk0 + (s - 1) * d.blk_step;
            const int pre = nkb < GRU_STAGES ? nkb : GRU_STAGES;
            if (lane == 0) {
                // W_hh tiles of the first stages do not depend on the previous step: request them before the step barrier
                for (int kb = 0; kb < pre; ++kb) {
                    const int st_i = (it + kb) % GRU_STAGES;
                    tc::mbar_wait(&empty_bar[st_i], (((it + kb) / GRU_STAGES) & 1) ^ 1);
                    uint8_t* st = smem + st_i * GRU_STAGE;
                    tc::mbar_arrive_expect_tx(&full_bar[st_i], GRU_STAGE);
                    uint8_t* wh = st + 2 * GRU_A_TILE;
                    uint8_t* wl = wh + GRU_W_TILE;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        tc::tma_load_2d(wh + g * GRU_U * 128, &mp.w_hi, &full_bar[st_i], kb * 64, g * H + j0);
                        tc::tma_load_2d(wl + g * GRU_U * 128, &mp.w_lo, &full_bar[st_i], kb * 64, g * H + j0);
                    }
                }
            }
            // step barrier: every CTA of this direction has published its h' of step s-1 (whole warp polls: 2 flags per lane)
            for (;;) {
                const unsigned fa = lane < nx ? ld_acquire_gpu(flags + lane) : (unsigned)s;
                const unsigned fb = lane + 32 < nx ? ld_acquire_gpu(flags + lane + 32) : (unsigned)s;
                if (__all_sync(0xffffffffu, fa >= (unsigned)s && fb >= (unsigned)s)) break;
                __nanosleep(100);
            }
            if (lane == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int st_i = it % GRU_STAGES;
                    uint8_t* st = smem + st_i * GRU_STAGE;
                    if (kb >= pre) {
                        tc::mbar_wait(&empty_bar[st_i], ((it / GRU_STAGES) & 1) ^ 1);
                        tc::mbar_arrive_expect_tx(&full_bar[st_i], GRU_STAGE);
                        uint8_t* wh = st + 2 * GRU_A_TILE;
                        uint8_t* wl = wh + GRU_W_TILE;
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            tc::tma_load_2d(wh + g * GRU_U * 128, &mp.w_hi, &full_bar[st_i], kb * 64, g * H + j0);
                            tc::tma_load_2d(wl + g * GRU_U * 128, &mp.w_lo, &full_bar[st_i], kb * 64, g * H + j0);
                        }
                    }
                    tc::tma_load_3d(st, &mp.h_hi, &full_bar[st_i], d.col0 + kb * 64, b0, blkp);
                    tc::tma_load_3d(st + GRU_A_TILE, &mp.h_lo, &full_bar[st_i], d.col0 + kb * 64, b0, blkp);
                }
            }
            it = __shfl_sync(0xffffffffu, it, 0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, GRU_N);
            uint32_t it = 0;
            for (int s = 1; s < d.nsteps; ++s) {
                tc::mbar_wait(tmem_empty_bar, ((s - 1) & 1) ^ 1);            // the epilogue has read the previous step's accumulator
                tc::tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int st_i = it % GRU_STAGES;
                    tc::mbar_wait(&full_bar[st_i], (it / GRU_STAGES) & 1);
                    tc::tc_fence_after();
                    const uint32_t st = tc::smem_u32(smem + st_i * GRU_STAGE);
                    const uint64_t a_hi = tc::umma_desc_sw128(st), a_lo = tc::umma_desc_sw128(st + GRU_A_TILE);
                    const uint64_t w_hi = tc::umma_desc_sw128(st + 2 * GRU_A_TILE), w_lo = tc::umma_desc_sw128(st + 2 * GRU_A_TILE + GRU_W_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                        tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                        tc::umma_bf16(tmem_base, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                    }
                    tc::umma_commit(&empty_bar[st_i]);
                }
                tc::umma_commit(tmem_full_bar);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = b0 + q * 32 + lane;
        const bool live = row < B;
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int s = 0; s < d.nsteps; ++s) {
            uint32_t ar[16], az[16], an[16];
            if (s > 0) {
                tc::mbar_wait(tmem_full_bar, (s - 1) & 1);
                tc::tc_fence_after();
                tc::tmem_ld_32x16(t0, ar);
                tc::tmem_ld_32x16(t0 + GRU_U, az);
                tc::tmem_ld_32x16(t0 + 2 * GRU_U, an);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tmem_empty_bar);
            } else {
#pragma unroll
                for (int u = 0; u < 16; ++u) ar[u] = az[u] = an[u] = 0u;      // h_0 = 0: W_hh h = 0
            }
            const int blk = d.blk0 + s * d.blk_step;
            if (live) {
                const float* gi = d.gi + (long long)s * d.gi_step + (size_t)row * d.ld_gi + j0;
                const size_t yo = ((size_t)blk * B + row) * d.ld_y + d.col0 + j0;
                const size_t yp = ((size_t)(blk - d.blk_step) * B + row) * d.ld_y + d.col0 + j0;
                float hn[16];
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    const float4 gr = ld4(gi + u), gz = ld4(gi + H + u), gn = ld4(gi + 2 * H + u);
                    const float4 hv = s > 0 ? ld4(d.y + yp + u) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 br = ld4(d.bhh + j0 + u), bz = ld4(d.bhh + H + j0 + u), bn = ld4(d.bhh + 2 * H + j0 + u);
                    const float grr[4] = {gr.x, gr.y, gr.z, gr.w}, gzz[4] = {gz.x, gz.y, gz.z, gz.w}, gnn[4] = {gn.x, gn.y, gn.z, gn.w};
                    const float hh[4] = {hv.x, hv.y, hv.z, hv.w}, brr[4] = {br.x, br.y, br.z, br.w}, bzz[4] = {bz.x, bz.y, bz.z, bz.w}, bnn[4] = {bn.x, bn.y, bn.z, bn.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float r = sigmoid_f(grr[i] + (__uint_as_float(ar[u + i]) + brr[i]));
                        const float z = sigmoid_f(gzz[i] + (__uint_as_float(az[u + i]) + bzz[i]));
                        const float n = tanhf(gnn[i] + r * (__uint_as_float(an[u + i]) + bnn[i]));
                        hn[u + i] = (1.0f - z) * n + z * hh[i];
                    }
                }
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    const float4 v = make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]);
                    st4(d.y + yo + u, v);
                    store_split4(d.ys, yo + u, v);
                    if (d.last_out && s == d.nsteps - 1) st4(d.last_out + (size_t)row * d.ld_last + j0 + u, v);
                }
            }
            // publish this CTA's h' of step s (the last step has no reader inside the kernel)
            if (s + 1 < d.nsteps) {
                asm volatile("bar.sync 1, 128;" ::: "memory");     // the 4 epilogue warps' stores are ordered before the release below
                if (threadIdx.x == 64) st_release_gpu(flags + blockIdx.x, (unsigned)(s + 1));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 64);
}
