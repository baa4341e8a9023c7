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

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % GRU_STAGES;
                const uint32_t ph = (kb / GRU_STAGES) & 1;
                tc::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * GRU_STAGE;
                tc::mbar_arrive_expect_tx(&full_bar[s], GRU_STAGE);
                tc::tma_load_2d(st, &mp.h_hi, &full_bar[s], kb * 64, b0);
                tc::tma_load_2d(st + GRU_A_TILE, &mp.h_lo, &full_bar[s], kb * 64, b0);
                uint8_t* wh = st + 2 * GRU_A_TILE;
                uint8_t* wl = wh + GRU_W_TILE;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    tc::tma_load_2d(wh + g * GRU_U * 128, &mp.w_hi, &full_bar[s], kb * 64, g * H + j0);
                    tc::tma_load_2d(wl + g * GRU_U * 128, &mp.w_lo, &full_bar[s], kb * 64, g * H + j0);
                }
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
