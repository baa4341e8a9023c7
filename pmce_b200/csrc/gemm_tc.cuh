// tcgen05 / TMA / TMEM GEMM for sm_100a with split-bf16 ("bf16x3") operands and a fused epilogue:
//     out[M,N] = epi( A[M,K] * W[N,K]^T ),   A = A_hi + A_lo,  W = W_hi + W_lo  (bf16 pairs, K-contiguous)
// Persistent: one CTA per SM loops over 128 x BN output tiles (n fastest, so concurrently running CTAs share the A
// tile through L2). Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..9 = epilogue (TMEM lane group = warp % 4, column half = (warp-2)/4; one output row per thread).
// Pipelines: smem ring of STAGES x {A_hi, A_lo, W_hi, W_lo} tiles ([rows][64 bf16], written by TMA with the 128-byte
// swizzle the UMMA descriptors expect) with full/empty mbarriers; TWO TMEM accumulators with tmem_full/tmem_empty
// mbarriers so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue (compile-time MODE), SIXTEEN warps: TMEM lane group = warp % 4, column group = (warp-2)/4 takes every fourth
// 16-column chunk: accumulator chunk (32 rows x 16 cols) TMEM -> registers -> (+bias, GELU, +residual) -> the warp's 2 KB
// shared-memory staging tile -> TMA store; the residual chunk arrives by TMA load into the same staging tile. Per chunk a
// warp waits for its previous store to have read the staging tile, so the time to drain an accumulator is set by how many
// such round trips are in flight: 16 warps x 16-column chunks keep twice as many in flight as 8 x 32 did (the epilogue,
// not the MMA stream, bounded the tile time: 52 us for the fc1 shape against 35 us of MMAs). The epilogue warps therefore issue no global loads/stores of their own: a row-per-thread direct store
// touches 32 different 128-byte lines per instruction and cost ~20 us per 128x256 tile (3x the MMA time) when measured.
// MODE TC_GENERIC keeps the direct path for strided ("mapped") outputs, row-embedding adds and unaligned N.
#pragma once
#include <stdlib.h>
#include "tc_common.cuh"
#include "common.cuh"
#include "host_once.h"

enum TcMode { TC_F32 = 0, TC_F32_RESID = 1, TC_SPLIT_GELU = 2, TC_GENERIC = 3, TC_NULL = 4, TC_SPLIT = 5, TC_ATTN32 = 6 };

// TC_ATTN32: the fused qkv projection of a 64-wide stream with 2 heads of 32 (N = 192 = q|k|v) writes the operand tiles of
// attn_rows_tc_kernel directly, as one bf16 tensor [M, 512] of 128-byte (row, head) records, so the attention kernel takes them
// by TMA in the layout its MMAs read (head_dim 32 = 64-byte operand rows: hi|lo halves are CONCATENATED into 128-byte rows):
//   cols [  0,128): QC = per head [q_hi | q_lo] (q already times softmax-scale * log2e)
//   cols [128,256): K1 = per head [k_hi | k_hi]
//   cols [256,384): K2 = per head [k_lo | k_lo]
//   cols [384,512): VC = per head [v_hi | v_lo]
constexpr int TC_ATT_LD = 512;

struct TcEpi {
    const float* bias;      // [N] or null
    const float* resid;     // fp32 [M, ld_resid] or null
    const float* rowadd;    // [period, N] or null (added at row % period)
    int rowadd_period;
    int ld_resid;
    int act;                // 0 none, 1 exact GELU, 2 ReLU (split-bf16 outputs and the generic path)
    float* out_f32;         // fp32 [M, ld_out] or null
    __nv_bfloat16* out_hi;  // split output [M, ld_split] or null (both hi and lo, or neither)
    __nv_bfloat16* out_lo;
    __nv_bfloat16* out_att; // TC_ATTN32: [M, TC_ATT_LD] operand records (N must be 192)
    int ld_out, ld_split;
    int mapped;             // out_f32 / resid are addressed as rmap(row) + cmap(col) instead of row*ld + col
    RowMap rmap, cmap;
    float qscale;           // TC_ATTN32: factor applied to the q columns before the split
    int dbg;                // profiling only (PMCE_TC_DBG): 1 = stage but do not issue TMA stores, 2 = no staging either
    int direct;             // epilogue stores straight from registers with 256-bit st.global (one full 32-byte sector per thread and
                            // instruction) instead of shared-memory staging + TMA store (PMCE_TC_DIRECT=1, A/B knob)
    int pair_relaxed;       // CTA pairs: release the accumulator with a relaxed cluster-scope arrive (PMCE_TC_PAIR_RELAXED=1)
    int wpre;               // request the W tiles of the ring's first pass before griddepcontrol.wait (PMCE_PDL_WPRE, default 1)
    int force_bn;           // 0 = tile width from the cost model; 32 / 64 / 128 / 256 = this launch's tile width (per-call A/B knobs)
};

struct TcOutMaps {          // TMA descriptors of the epilogue tensors (box = 32 rows x 16 columns)
    CUtensorMap out;        // fp32 out (SWIZZLE_64B) or bf16 hi (no swizzle, 32-byte rows)
    CUtensorMap out_lo;     // bf16 lo
    CUtensorMap resid;      // fp32 residual (SWIZZLE_64B)
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;            // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // TMA warp + MMA warp + epilogue warps
constexpr int TC_CW = 16;            // epilogue chunk width in columns
constexpr int TC_STG = 2048;         // staging bytes per epilogue warp: [32][16] fp32, or [32][16] bf16 hi + lo

// NBUF: staging tiles per epilogue warp. A warp may only refill a staging tile once the TMA store that reads it has drained; with
// one tile per warp the epilogue's pace is (staging bytes in flight) / (store drain latency), with two the warp fills tile b while
// the store of tile b^1 drains (cp.async.bulk.wait_group.read 1).
template <int BN, int NCTA = 1, int NBUF = 1>
struct TcCfg {
    static constexpr int A_TILE = TC_BM * 128;          // bytes per A half (hi or lo)
    static constexpr int W_TILE = (BN / NCTA) * 128;    // per CTA: a CTA pair holds half of the tile's W rows each
    static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * W_TILE;
    static constexpr int STG_BYTES = TC_EPI_WARPS * TC_STG * NBUF;   // per-warp epilogue staging tiles (1024-B aligned)
    static constexpr int RING_BYTES = 227 * 1024 - 1024 - 256 - STG_BYTES;
    static constexpr int STAGES = RING_BYTES / STAGE_BYTES >= 4 ? 4 : RING_BYTES / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // two accumulators
    static_assert(STAGES >= 2, "tile too large");
    static_assert(TMEM_COLS <= 512, "TMEM");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
    static_assert(NCTA == 1 || (NCTA == 2 && BN == 256), "CTA pairs run 256 x 256 tiles");
};

// Direct (row-per-thread) epilogue for one 16-column chunk: runtime flags, any addressing. Used by TC_GENERIC.
__device__ __forceinline__ void tc_epilogue_generic(const TcEpi& e, const uint32_t (&v)[TC_CW], int row, int col0, int N) {
    const float* radd = e.rowadd ? e.rowadd + (size_t)(row % e.rowadd_period) * N : nullptr;
#pragma unroll 1
    for (int i = 0; i < TC_CW; ++i) {
        const int col = col0 + i;
        if (col >= N) break;
        float x = __uint_as_float(v[i]);
        if (e.bias) x += e.bias[col];
        if (e.act == 1) x = gelu_erf(x);
        if (e.act == 2) x = fmaxf(x, 0.f);
        if (radd) x += radd[col];
        if (e.mapped) {
            const long long off = e.rmap(row) + e.cmap(col);
            if (e.resid) x += e.resid[off];
            if (e.out_f32) e.out_f32[off] = x;
        } else {
            if (e.resid) x += e.resid[(size_t)row * e.ld_resid + col];
            if (e.out_f32) e.out_f32[(size_t)row * e.ld_out + col] = x;
        }
        if (e.out_hi) {
            __nv_bfloat16 h, l;
            tc::split_bf16(x, h, l);
            e.out_hi[(size_t)row * e.ld_split + col] = h;
            e.out_lo[(size_t)row * e.ld_split + col] = l;
        }
    }
}

// NCTA == 2: the same kernel on CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256 x 256 tile per pair, each CTA loads
// its 128 rows of A and HALF of the tile's W rows (a third less operand traffic L2->SM and smem per FLOP than two 128 x 256
// tiles), the leader issues the MMAs for both, each CTA drains its own 128 accumulator rows.
// CL > 1 (with NCTA == 1): a cluster of CL CTAs works on CL consecutive 128-row m-tiles of the SAME n-tile; every CTA loads its own
// A rows and ONE CL-th of the W tile, multicast to all CTAs of the cluster - W leaves L2 once per cluster instead of once per CTA
// (the GEMMs are bound by L2->SMEM operand bytes): 48 KB per CTA and k-block instead of 64 KB on a CTA pair.
template <int BN, int MODE, int NCTA, int NBUF = 1, int CL = 1>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const __grid_constant__ TcOutMaps om, int M, int N, int K, TcEpi e) {
    using Cfg = TcCfg<BN, NCTA, NBUF>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stg_all = smem + STAGES * Cfg::STAGE_BYTES;                 // 1024-B aligned (stage sizes are multiples of 1 KB)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_all + Cfg::STG_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint64_t* resid_bar = tmem_empty_bar + 2;          // [TC_EPI_WARPS]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(resid_bar + TC_EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (K + TC_BK - 1) / TC_BK;
    static_assert(CL == 1 || (NCTA == 1 && BN % CL == 0), "W multicast clusters are built from single-CTA tiles");
    const int tiles_n = (N + BN - 1) / BN, tiles_m = ((M + NCTA * TC_BM - 1) / (NCTA * TC_BM) + CL - 1) / CL;   // CL > 1: groups of CL m-tiles
    const int num_tiles = tiles_n * tiles_m;
    const uint32_t crank = (NCTA == 2 || CL > 1) ? tc::cluster_ctarank() : 0;
    const uint32_t rank = NCTA == 2 ? crank : 0;                          // position in the CTA pair
    const int tile0 = blockIdx.x / (NCTA * CL), tile_step = gridDim.x / (NCTA * CL);    // tiles are dealt to pairs / clusters
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA_hi); tc::tma_prefetch_desc(&tmA_lo);
        tc::tma_prefetch_desc(&tmW_hi); tc::tma_prefetch_desc(&tmW_lo);
        if (MODE == TC_F32 || MODE == TC_F32_RESID || MODE == TC_SPLIT_GELU || MODE == TC_SPLIT || MODE == TC_ATTN32) tc::tma_prefetch_desc(&om.out);
        if (MODE == TC_SPLIT_GELU || MODE == TC_SPLIT) tc::tma_prefetch_desc(&om.out_lo);
        if (MODE == TC_F32_RESID) tc::tma_prefetch_desc(&om.resid);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], CL); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], NCTA * TC_EPI_WARPS); }
        for (int w = 0; w < TC_EPI_WARPS; ++w) tc::mbar_init(&resid_bar[w], 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
    }
    if (warp == 1) { if (NCTA == 2) tc::tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS); else tc::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); }
    tc::tc_fence_before();
    __syncthreads();
    if (NCTA == 2 || CL > 1) tc::cluster_sync_all();     // the peers' barriers are initialised before anything signals them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();      // TMEM is ours: the next kernel's CTAs may be scheduled behind this one
    // everything above overlapped the previous kernel's tail; its outputs (A, residual) are visible after the wait. The producer
    // thread waits later: it first requests the W tiles (constants) of the ring's first pass
    if (threadIdx.x != 0) pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // first pass over the ring (every stage is free): W tiles before the grid dependency resolves, A tiles after
            uint32_t pre = 0;
            if (CL == 1 && e.wpre) {
                for (int tile = tile0; tile < num_tiles && pre < (uint32_t)STAGES; tile += tile_step) {
                    const int n0 = (tile % tiles_n) * BN + (int)rank * (BN / NCTA);
                    for (int kb = 0; kb < nkb && pre < (uint32_t)STAGES; ++kb, ++pre) {
                        uint8_t* st = smem + pre * Cfg::STAGE_BYTES;
                        if (NCTA == 1) {
                            tc::mbar_arrive_expect_tx(&full_bar[pre], Cfg::STAGE_BYTES);
                            tc::tma_load_2d(st + 2 * Cfg::A_TILE, &tmW_hi, &full_bar[pre], kb * TC_BK, n0);
                            tc::tma_load_2d(st + 2 * Cfg::A_TILE + Cfg::W_TILE, &tmW_lo, &full_bar[pre], kb * TC_BK, n0);
                        } else {
                            const uint32_t fb = tc::mapa_rank(tc::smem_u32(&full_bar[pre]), 0);
                            if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[pre], 2 * Cfg::STAGE_BYTES);
                            tc::tma_load_2d_pair(st + 2 * Cfg::A_TILE, &tmW_hi, fb, kb * TC_BK, n0);
                            tc::tma_load_2d_pair(st + 2 * Cfg::A_TILE + Cfg::W_TILE, &tmW_lo, fb, kb * TC_BK, n0);
                        }
                    }
                }
            }
            pdl_wait();
            uint32_t it = 0;                                    // global k-block counter across tiles
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int m0 = ((tile / tiles_n) * CL + (CL > 1 ? (int)crank : 0)) * (NCTA * TC_BM) + (int)rank * TC_BM,
                          n0 = (tile % tiles_n) * BN + (int)rank * (BN / NCTA);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    if (it < pre) {                             // stage armed and its W tiles requested above
                        if (NCTA == 1) {
                            tc::tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                            tc::tma_load_2d(st + Cfg::A_TILE, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                        } else {
                            const uint32_t fb = tc::mapa_rank(tc::smem_u32(&full_bar[s]), 0);
                            tc::tma_load_2d_pair(st, &tmA_hi, fb, kb * TC_BK, m0);
                            tc::tma_load_2d_pair(st + Cfg::A_TILE, &tmA_lo, fb, kb * TC_BK, m0);
                        }
                        continue;
                    }
                    tc::mbar_wait(&empty_bar[s], ph ^ 1);
                    if (CL > 1) {
                        // own A rows; this CTA's CL-th of the W tile goes to every CTA of the cluster (all CTAs' MMAs have released the stage)
                        constexpr int WQ = Cfg::W_TILE / CL;
                        tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                        tc::tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                        tc::tma_load_2d(st + Cfg::A_TILE, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                        tc::tma_load_2d_mcast(st + 2 * Cfg::A_TILE + crank * WQ, &tmW_hi, &full_bar[s], kb * TC_BK, n0 + (int)crank * (BN / CL), CMASK);
                        tc::tma_load_2d_mcast(st + 2 * Cfg::A_TILE + Cfg::W_TILE + crank * WQ, &tmW_lo, &full_bar[s], kb * TC_BK, n0 + (int)crank * (BN / CL), CMASK);
                    } else if (NCTA == 1) {
                        tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                        tc::tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                        tc::tma_load_2d(st + Cfg::A_TILE, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                        tc::tma_load_2d(st + 2 * Cfg::A_TILE, &tmW_hi, &full_bar[s], kb * TC_BK, n0);
                        tc::tma_load_2d(st + 2 * Cfg::A_TILE + Cfg::W_TILE, &tmW_lo, &full_bar[s], kb * TC_BK, n0);
                    } else {
                        // both CTAs' bytes are counted on the LEADER's barrier (the only MMA issuer waits there)
                        const uint32_t fb = tc::mapa_rank(tc::smem_u32(&full_bar[s]), 0);
                        if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
                        tc::tma_load_2d_pair(st, &tmA_hi, fb, kb * TC_BK, m0);
                        tc::tma_load_2d_pair(st + Cfg::A_TILE, &tmA_lo, fb, kb * TC_BK, m0);
                        tc::tma_load_2d_pair(st + 2 * Cfg::A_TILE, &tmW_hi, fb, kb * TC_BK, n0);
                        tc::tma_load_2d_pair(st + 2 * Cfg::A_TILE + Cfg::W_TILE, &tmW_lo, fb, kb * TC_BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(NCTA * TC_BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tcount) {
                const uint32_t acc = tcount & 1;
                tc::mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);     // epilogue(s) have drained this accumulator
                tc::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    tc::mbar_wait(&full_bar[s], ph);
                    tc::tc_fence_after();
                    const uint32_t st = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = tc::umma_desc_sw128(st), a_lo = tc::umma_desc_sw128(st + Cfg::A_TILE);
                    const uint64_t w_hi = tc::umma_desc_sw128(st + 2 * Cfg::A_TILE), w_lo = tc::umma_desc_sw128(st + 2 * Cfg::A_TILE + Cfg::W_TILE);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        // small cross terms first, then the leading term
                        if (NCTA == 1) {
                            tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                            tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                            tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                        } else {
                            tc::umma_bf16_pair(tmem_d, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                            tc::umma_bf16_pair(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                            tc::umma_bf16_pair(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                        }
                    }
                    if (CL > 1) tc::umma_commit_mcast(&empty_bar[s], CMASK);                       // every CTA of the cluster may refill its share of the stage
                    else if (NCTA == 1) tc::umma_commit(&empty_bar[s]); else tc::umma_commit_pair(&empty_bar[s]);   // smem slot free once these MMAs have read it
                }
                if (NCTA == 1) tc::umma_commit(&tmem_full_bar[acc]); else tc::umma_commit_pair(&tmem_full_bar[acc]);   // accumulator complete
            }
        }
    } else {
        // ---- epilogue: 16 warps; lane group q = warp % 4 (TMEM lanes [32q,32q+32) = tile rows), column group cg = (warp-2)/4
        //      takes the 16-column chunks cg, cg+4, cg+8, ... ----
        const int ew = warp - 2;
        const int q = warp & 3;
        const int cg = ew >> 2;
        constexpr int CH = BN / TC_CW;                        // 16-column chunks per tile
        uint8_t* const stg_base = stg_all + ew * TC_STG * NBUF;   // this warp's staging tile(s)
        uint32_t tcount = 0, rcount = 0, ccount = 0;
        bool store_pending = false;
        // accumulator-drained arrivals go to the leader's barrier (its MMA thread is the only waiter)
        const uint32_t te_bar[2] = {tc::mapa_rank(tc::smem_u32(&tmem_empty_bar[0]), 0), tc::mapa_rank(tc::smem_u32(&tmem_empty_bar[1]), 0)};
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tcount) {
            const int m0 = ((tile / tiles_n) * CL + (CL > 1 ? (int)crank : 0)) * (NCTA * TC_BM) + (int)rank * TC_BM, n0 = (tile % tiles_n) * BN;
            const uint32_t acc = tcount & 1;
            const int row0 = m0 + q * 32;
            const int row = row0 + lane;
            tc::mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
            tc::tc_fence_after();
            if (cg >= CH) {                                       // BN == 32: column groups 2 and 3 have no chunk
                tc::tc_fence_before();
                if (lane == 0) {
                    if (NCTA == 1) tc::mbar_arrive(&tmem_empty_bar[acc]);
                    else if (e.pair_relaxed) tc::mbar_arrive_cluster_relaxed(te_bar[acc]);
                    else tc::mbar_arrive_cluster(te_bar[acc]);
                }
                continue;
            }
            const int c_last = cg + 4 * ((CH - 1 - cg) / 4);
#pragma unroll 1
            for (int c = cg; c < CH; c += 4) {
                const int col0 = n0 + c * TC_CW;
                const bool live = row0 < M && col0 < N;          // warp-uniform
                uint8_t* const stg = stg_base + (NBUF > 1 ? (ccount % NBUF) * TC_STG : 0);
                const uint32_t stg_u32 = tc::smem_u32(stg);
                if (live) ++ccount;
                if (MODE == TC_F32_RESID && live) {
                    // the staging tile is about to be overwritten by the residual load: the previous store must have read it
                    if (lane == 0) {
                        if (store_pending) tc::tma_store_wait_read<NBUF - 1>();
                        tc::mbar_arrive_expect_tx(&resid_bar[ew], TC_STG);
                        tc::tma_load_2d(stg, &om.resid, &resid_bar[ew], col0, row0);
                    }
                    store_pending = false;
                }
                uint32_t v[TC_CW];
                tc::tmem_ld_32x16(tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * TC_CW), v);
                tc::tmem_ld_wait();
                if (c == c_last) {                               // last read of this accumulator by this warp: release it early
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (NCTA == 1) tc::mbar_arrive(&tmem_empty_bar[acc]);
                        else if (e.pair_relaxed) tc::mbar_arrive_cluster_relaxed(te_bar[acc]);
                        else tc::mbar_arrive_cluster(te_bar[acc]);
                    }
                }
                if (!live) continue;                             // warp-uniform
                if (MODE == TC_GENERIC) {
                    if (row < M) tc_epilogue_generic(e, v, row, col0, N);
                    __syncwarp();
                } else if (MODE == TC_F32 || MODE == TC_F32_RESID) {
                    float f[TC_CW];
#pragma unroll
                    for (int i = 0; i < TC_CW; i += 4) {
                        const float4 b = ld4(e.bias + col0 + i);     // N % 16 == 0 in the TMA modes (checked by the launcher)
                        f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
                        f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
                    }
                    // [32 rows][16 fp32] tile, 64-byte rows, SWIZZLE_64B: 16-B chunk ^= (row >> 1) & 3
                    const uint32_t rowaddr = stg_u32 + lane * 64;
                    const int sw = (lane >> 1) & 3;
                    if (MODE == TC_F32_RESID) {
                        tc::mbar_wait(&resid_bar[ew], rcount & 1);
                        ++rcount;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 r;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(rowaddr + ((j ^ sw) << 4)));
                            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
                        }
                    } else {
                        if (store_pending) { if (lane == 0) tc::tma_store_wait_read<NBUF - 1>(); __syncwarp(); }
                    }
                    if (e.dbg == 2) { if (f[0] == 123.456f) e.out_f32[0] = f[1]; continue; }
                    if (MODE == TC_F32 && e.direct) {
                        if (row < M) {
                            float* o = e.out_f32 + (size_t)row * e.ld_out + col0;
                            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]),
                                         "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
                            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 8), "f"(f[8]), "f"(f[9]), "f"(f[10]), "f"(f[11]),
                                         "f"(f[12]), "f"(f[13]), "f"(f[14]), "f"(f[15]) : "memory");
                        }
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + ((j ^ sw) << 4)), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                     "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (e.dbg == 1) continue;
                    if (lane == 0) { tc::tma_store_2d(&om.out, stg, col0, row0); tc::tma_store_commit(); }
                    store_pending = true;
                } else if (MODE == TC_SPLIT_GELU || MODE == TC_SPLIT || MODE == TC_ATTN32) {
                    uint32_t hi[TC_CW / 2], lo[TC_CW / 2];
                    const float qs = (MODE == TC_ATTN32 && col0 < 64) ? e.qscale : 1.0f;
#pragma unroll
                    for (int i = 0; i < TC_CW; i += 4) {
                        const float4 b = ld4(e.bias + col0 + i);
                        float x0 = __uint_as_float(v[i]) + b.x, x1 = __uint_as_float(v[i + 1]) + b.y;
                        float x2 = __uint_as_float(v[i + 2]) + b.z, x3 = __uint_as_float(v[i + 3]) + b.w;
                        if (MODE == TC_SPLIT_GELU) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3); }
                        if (MODE == TC_SPLIT && e.act == 2) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }   // ReLU
                        if (MODE == TC_ATTN32) { x0 *= qs; x1 *= qs; x2 *= qs; x3 *= qs; }
                        tc::split_bf16x2(x0, x1, hi[i / 2], lo[i / 2]);
                        tc::split_bf16x2(x2, x3, hi[i / 2 + 1], lo[i / 2 + 1]);
                    }
                    if ((MODE == TC_SPLIT_GELU || MODE == TC_SPLIT) && e.direct) {
                        if (row < M) {
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(e.out_hi + (size_t)row * e.ld_split + col0), "r"(hi[0]),
                                         "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(e.out_lo + (size_t)row * e.ld_split + col0), "r"(lo[0]),
                                         "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
                        }
                        continue;
                    }
                    if (store_pending) { if (lane == 0) tc::tma_store_wait_read<NBUF - 1>(); __syncwarp(); }
                    // two dense [32 rows][16 bf16] tiles (32-byte rows, no swizzle): hi at +0, lo at +1024
                    const uint32_t rowaddr = stg_u32 + lane * 32;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + (j << 4)), "r"(hi[4 * j]), "r"(hi[4 * j + 1]),
                                     "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + 1024 + (j << 4)), "r"(lo[4 * j]), "r"(lo[4 * j + 1]),
                                     "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3]) : "memory");
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (MODE == TC_ATTN32) {
                            // col0 = 64 * which + 32 * head + d0 (d0 = 0 or 16) -> record column 64 * head + d0 (+ 32 for the second half)
                            const int which = col0 >> 6, rc = ((col0 >> 5) & 1) * 64 + (col0 & 31);
                            if (which == 0) {            // q: [hi | lo]
                                tc::tma_store_2d(&om.out, stg, rc, row0);
                                tc::tma_store_2d(&om.out, stg + 1024, rc + 32, row0);
                            } else if (which == 1) {     // k: K1 = [hi | hi], K2 = [lo | lo]
                                tc::tma_store_2d(&om.out, stg, 128 + rc, row0);
                                tc::tma_store_2d(&om.out, stg, 128 + rc + 32, row0);
                                tc::tma_store_2d(&om.out, stg + 1024, 256 + rc, row0);
                                tc::tma_store_2d(&om.out, stg + 1024, 256 + rc + 32, row0);
                            } else {                     // v: [hi | lo]
                                tc::tma_store_2d(&om.out, stg, 384 + rc, row0);
                                tc::tma_store_2d(&om.out, stg + 1024, 384 + rc + 32, row0);
                            }
                        } else {
                            tc::tma_store_2d(&om.out, stg, col0, row0);
                            tc::tma_store_2d(&om.out_lo, stg + 1024, col0, row0);
                        }
                        tc::tma_store_commit();
                    }
                    store_pending = true;
                }
            }
        }
        if (lane == 0) tc::tma_store_wait_read<0>();         // the staging tile must outlive the reads; the grid boundary orders the writes
    }
    tc::tc_fence_before();
    __syncthreads();
    if (NCTA == 2 || CL > 1) tc::cluster_sync_all();     // no CTA leaves (or frees TMEM) while a peer may still touch it
    if (warp == 1) { if (NCTA == 2) tc::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// fp32 [rows, cols] (ld) -> bf16 hi / lo [rows, ld_out]; optional relu on the way (linear_cur input).
__global__ void split_rows_kernel(const float* __restrict__ x, int rows, int cols, int ld, int relu, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ld_out) {
    pdl_enter();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per 4 elements
    const int c4 = cols / 4;
    if (idx >= (size_t)rows * c4) return;
    const int r = (int)(idx / c4), c = (int)(idx % c4) * 4;
    float4 v = ld4(x + (size_t)r * ld + c);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    uint2 h, l;
    tc::split_bf16x2(v.x, v.y, h.x, l.x);
    tc::split_bf16x2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + (size_t)r * ld_out + c) = h;
    *reinterpret_cast<uint2*>(lo + (size_t)r * ld_out + c) = l;
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// row-major [rows, cols] matrix with row stride ld (elements); box = [box_rows][box_cols]; zero OOB fill / clipped stores.
static inline int make_tmap(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int rows, int cols, int ld, int box_cols,
                            int box_rows, CUtensorMapSwizzle sw) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}
// bf16 3-D view [d2][d1][cols] with element strides (ld1, ld2) of one K-major operand tensor; box = [b2][b1][64], SW128
static inline int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, int cols, int d1, int d2, long long ld1, long long ld2, int b1, int b2) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)ld1 * 2, (cuuint64_t)ld2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}
// bf16 operand tile map: box = [box_rows][64], 128-byte swizzle
static inline int make_tmap_bf16(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows) {
    return make_tmap(m, ptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, cols, ld, TC_BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

struct TcOperand {   // split bf16 matrix [rows, cols], row stride ld (elements)
    const __nv_bfloat16* hi; const __nv_bfloat16* lo; int rows, cols, ld;
};

// 256-wide tiles run on CTA pairs (cta_group::2) when the problem is large in N and K: measured with tools/gemm_sweep.py,
// M=17408: N=1536,K=512 69.8 -> 63.4 us, N=1024,K=512 (GELU) 55.5 -> 52.3, N=512,K=1024 48.5 -> 45.9, N=512,K=512 unchanged;
// K=256 shapes lose 1-2 us (three 64 KB stages hold fewer k-blocks than the tile needs to hide the fill).
// PMCE_TC_PAIR=0 disables, =2 forces pairs for every 256-wide tile (tests).
static inline bool tc_pair_enabled(int M, int N, int K) {
    static int mode = -1;
    if (mode < 0) mode = pmce_env_int("PMCE_TC_PAIR", 1);
    if (mode == 0 || M <= TC_BM) return false;
    return mode >= 2 || (N >= 512 && K >= 512);
}

constexpr int TC_CL = 4;      // CTAs per W-multicast cluster
// PMCE_TC_MCAST=1: 256-wide tiles of large GEMMs run on clusters of 4 single-CTA tiles with the W tile multicast (instead of CTA pairs)
static inline bool tc_mcast_enabled(int M, int N, int K) {
    static int mode = -1;
    if (mode < 0) mode = pmce_env_int("PMCE_TC_MCAST", 0);
    return mode != 0 && M > TC_CL * TC_BM && N >= 512 && K >= 512;
}

// SMs the persistent grids leave alone. pmce_forward sets it while the image-feature stream's few-CTA GRU steps run beside the
// pose lifter (api.cu): the GEMM CTAs then never wait behind a GRU CTA and vice versa. Host-side, per calling thread.
static thread_local int tc_sm_reserve = 0;

// Persistent grid for `tiles` equal work items dealt round-robin to at most max_units CTAs (pairs): the SMALLEST grid with the
// same number of waves. 272 tiles on 74 pairs take 4 waves, so do 68 pairs - the 6 TPCs left over cost nothing and are free for
// a concurrent stream.
static inline int tc_balanced_grid(long long tiles, int max_units) {
    if (max_units < 1) max_units = 1;
    if (tiles <= max_units) return (int)tiles;
    const long long waves = (tiles + max_units - 1) / max_units;
    return (int)((tiles + waves - 1) / waves);
}

template <int BN, int MODE>
static inline int launch_linear_tc_mode(const CUtensorMap* ta, const CUtensorMap* tw, const TcOutMaps& om, int M, int N, int K, const TcEpi& e,
                                        cudaStream_t st) {
    if (BN == 256 && MODE != TC_GENERIC && MODE != TC_NULL && tc_mcast_enabled(M, N, K)) {
        constexpr int BNP = BN == 256 ? 256 : 256;
        using K4 = TcCfg<BNP, 1, 1>;
        if (!pmce_configure_smem<linear_tc_kernel<BNP, MODE, 1, 1, TC_CL>>(K4::SMEM_BYTES)) return 2;
        const long long tiles = (long long)((N + 255) / 256) * (((M + TC_BM - 1) / TC_BM + TC_CL - 1) / TC_CL);
        const int clusters = (int)(tiles < tc_num_sms() / TC_CL ? tiles : tc_num_sms() / TC_CL);
        return pmce_launch(linear_tc_kernel<BNP, MODE, 1, 1, TC_CL>, dim3(TC_CL * clusters), dim3(TC_THREADS), K4::SMEM_BYTES, st, TC_CL,
                           ta[0], ta[1], tw[4], tw[5], om, M, N, K, e) == cudaSuccess ? 0 : 3;
    }
    if (BN == 256 && tc_pair_enabled(M, N, K)) {
        constexpr int BNP = BN == 256 ? 256 : 256;
        static int nbuf = -1;     // PMCE_TC_NBUF=2: two staging tiles per epilogue warp and a two-stage operand ring (A/B knob)
        if (nbuf < 0) nbuf = pmce_env_int("PMCE_TC_NBUF", 1);
        const long long tiles = (long long)((N + 255) / 256) * ((M + 2 * TC_BM - 1) / (2 * TC_BM));
        const int pairs = tc_balanced_grid(tiles, (tc_num_sms() - tc_sm_reserve) / 2);
        if (nbuf == 2 && MODE != TC_GENERIC && MODE != TC_NULL) {
            if (!pmce_configure_smem<linear_tc_kernel<BNP, MODE, 2, 2>>(TcCfg<BNP, 2, 2>::SMEM_BYTES)) return 2;
            return pmce_launch(linear_tc_kernel<BNP, MODE, 2, 2>, dim3(2 * pairs), dim3(TC_THREADS), TcCfg<BNP, 2, 2>::SMEM_BYTES, st, 2,
                               ta[0], ta[1], tw[2], tw[3], om, M, N, K, e) == cudaSuccess ? 0 : 3;
        }
        if (!pmce_configure_smem<linear_tc_kernel<BNP, MODE, 2>>(TcCfg<BNP, 2>::SMEM_BYTES)) return 2;
        return pmce_launch(linear_tc_kernel<BNP, MODE, 2>, dim3(2 * pairs), dim3(TC_THREADS), TcCfg<BNP, 2>::SMEM_BYTES, st, 2,
                           ta[0], ta[1], tw[2], tw[3], om, M, N, K, e) == cudaSuccess ? 0 : 3;
    }
    if (!pmce_configure_smem<linear_tc_kernel<BN, MODE, 1>>(TcCfg<BN>::SMEM_BYTES)) return 2;
    const long long tiles = (long long)((N + BN - 1) / BN) * ((M + TC_BM - 1) / TC_BM);
    const int grid = tc_balanced_grid(tiles, tc_num_sms() - tc_sm_reserve);
    return pmce_launch(linear_tc_kernel<BN, MODE, 1>, dim3(grid), dim3(TC_THREADS), TcCfg<BN>::SMEM_BYTES, st, 0,
                       ta[0], ta[1], tw[0], tw[1], om, M, N, K, e) == cudaSuccess ? 0 : 3;
}

template <int BN>
static inline int launch_linear_tc_bn(const TcOperand& A, const TcOperand& W, const TcEpi& e, cudaStream_t st) {
    // W maps: [0,1] whole-tile box, [2,3] half (a CTA of a pair loads half of the W rows), [4,5] quarter (a CTA of a multicast
    // cluster loads a quarter); launch_linear_tc_mode picks the pair that matches the kernel variant it launches
    CUtensorMap ta[2], tw[6];
    if (make_tmap_bf16(&ta[0], A.hi, A.rows, A.cols, A.ld, TC_BM) || make_tmap_bf16(&ta[1], A.lo, A.rows, A.cols, A.ld, TC_BM) ||
        make_tmap_bf16(&tw[0], W.hi, W.rows, W.cols, W.ld, BN) || make_tmap_bf16(&tw[1], W.lo, W.rows, W.cols, W.ld, BN))
        return 1;
    if (BN == 256 && (make_tmap_bf16(&tw[2], W.hi, W.rows, W.cols, W.ld, BN / 2) || make_tmap_bf16(&tw[3], W.lo, W.rows, W.cols, W.ld, BN / 2) ||
                      make_tmap_bf16(&tw[4], W.hi, W.rows, W.cols, W.ld, BN / TC_CL) || make_tmap_bf16(&tw[5], W.lo, W.rows, W.cols, W.ld, BN / TC_CL)))
        return 1;
    const int M = A.rows, N = W.rows, K = A.cols;
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    static int null_epi = -1;   // PMCE_TC_NULL=1: skip the epilogue entirely (mainloop-only timing, tools/gemm_sweep.py)
    if (null_epi < 0) null_epi = pmce_profiling_knob("PMCE_TC_NULL") ? 1 : 0;
    static int dbg = -1;
    if (dbg < 0) dbg = pmce_profiling_knob("PMCE_TC_DBG");
    const_cast<TcEpi&>(e).dbg = dbg;
    static int relaxed = -1;   // ncu shows the release.cluster arrive (a cluster-scope fence per accumulator release) at 15 % of the pair kernel's stall samples
    if (relaxed < 0) relaxed = pmce_env_int("PMCE_TC_PAIR_RELAXED", 0) ? 1 : 0;
    const_cast<TcEpi&>(e).pair_relaxed = relaxed;
    const_cast<TcEpi&>(e).wpre = pmce_env_int("PMCE_PDL_WPRE", 1) ? 1 : 0;      // live, like PMCE_PDL
    static int direct = -1;
    if (direct < 0) direct = pmce_env_int("PMCE_TC_DIRECT", 0);
    // 256-bit stores need 32-byte aligned rows: leading dimensions in multiples of 16 bf16 / 8 fp32 and 32-byte aligned bases
    auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    const_cast<TcEpi&>(e).direct = (direct && ((e.out_hi && a32(e.out_hi) && a32(e.out_lo) && e.ld_split % 16 == 0) ||
                                               (e.out_f32 && !e.resid && a32(e.out_f32) && e.ld_out % 8 == 0))) ? 1 : 0;
    TcOutMaps om;
    memset(&om, 0, sizeof(om));
    if (null_epi) return launch_linear_tc_mode<BN, TC_NULL>(ta, tw, om, M, N, K, e, st);
    // TMA epilogues need: plain row-major addressing, bias present, N % 16 == 0, 16-byte aligned bases and row strides
    const bool plain = !e.mapped && !e.rowadd && e.bias && a16(e.bias) && (N % TC_CW == 0);
    if (plain && e.out_f32 && !e.out_hi && e.act == 0 && a16(e.out_f32) && e.ld_out % 4 == 0) {
        if (make_tmap(&om.out, e.out_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, e.ld_out, TC_CW, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
        if (!e.resid) return launch_linear_tc_mode<BN, TC_F32>(ta, tw, om, M, N, K, e, st);
        if (a16(e.resid) && e.ld_resid % 4 == 0) {
            if (make_tmap(&om.resid, e.resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, e.ld_resid, TC_CW, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
            return launch_linear_tc_mode<BN, TC_F32_RESID>(ta, tw, om, M, N, K, e, st);
        }
    }
    if (e.out_att) {
        if (!plain || N != 192 || e.out_f32 || e.out_hi || e.resid || e.act || !a16(e.out_att)) return 4;
        if (make_tmap(&om.out, e.out_att, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, TC_ATT_LD, TC_ATT_LD, TC_CW, 32, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
        return launch_linear_tc_mode<BN, TC_ATTN32>(ta, tw, om, M, N, K, e, st);
    }
    if (plain && e.out_hi && !e.out_f32 && !e.resid && a16(e.out_hi) && a16(e.out_lo) && e.ld_split % 8 == 0) {
        if (make_tmap(&om.out, e.out_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, e.ld_split, TC_CW, 32, CU_TENSOR_MAP_SWIZZLE_NONE) ||
            make_tmap(&om.out_lo, e.out_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, e.ld_split, TC_CW, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
            return 1;
        if (e.act == 1) return launch_linear_tc_mode<BN, TC_SPLIT_GELU>(ta, tw, om, M, N, K, e, st);
        return launch_linear_tc_mode<BN, TC_SPLIT>(ta, tw, om, M, N, K, e, st);
    }
    return launch_linear_tc_mode<BN, TC_GENERIC>(ta, tw, om, M, N, K, e, st);
}

// Tile width: the BN in {256,128,64,32} with the lowest estimated launch time (rounds x per-round cost).
// Requirements: A.cols == W.cols (K), K % 8 == 0, ld % 8 == 0, 16-byte aligned bases.
static inline int launch_linear_tc(const TcOperand& A, const TcOperand& W, const TcEpi& e, cudaStream_t st) {
    const int N = W.rows;
    const long long tiles_m = (A.rows + TC_BM - 1) / TC_BM;
    const int sms = tc_num_sms();
    static int forced = -1;   // PMCE_TC_BN=<32|64|128|256>: tile-sweep knob for profiling (tools/gemm_sweep.py)
    if (forced < 0) forced = pmce_env_int("PMCE_TC_BN", 0);
    if (e.force_bn == 256) return launch_linear_tc_bn<256>(A, W, e, st);
    if (e.force_bn == 128) return launch_linear_tc_bn<128>(A, W, e, st);
    if (e.force_bn == 64) return launch_linear_tc_bn<64>(A, W, e, st);
    if (e.force_bn == 32) return launch_linear_tc_bn<32>(A, W, e, st);
    if (forced == 256) return launch_linear_tc_bn<256>(A, W, e, st);
    if (forced == 128) return launch_linear_tc_bn<128>(A, W, e, st);
    if (forced == 64) return launch_linear_tc_bn<64>(A, W, e, st);
    if (forced == 32) return launch_linear_tc_bn<32>(A, W, e, st);
    // cost model fitted to tools/gemm_sweep.py: a launch takes ceil(tiles / SMs) rounds and a round costs a fixed part
    // (pipeline fill, one epilogue drain) plus a part proportional to the tile width; ties go to the wider tile (fewer A re-reads)
    int best = 32;
    long long best_cost = -1;
    for (int bn = 256; bn >= 32; bn >>= 1) {
        if (bn > 32 && bn / 2 >= N) continue;                     // at least half of the tile's columns must exist
        const long long tiles = tiles_m * ((N + bn - 1) / bn);
        const long long cost = ((tiles + sms - 1) / sms) * (32 + bn);
        if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
    }
    if (best == 256) return launch_linear_tc_bn<256>(A, W, e, st);
    if (best == 128) return launch_linear_tc_bn<128>(A, W, e, st);
    if (best == 64) return launch_linear_tc_bn<64>(A, W, e, st);
    return launch_linear_tc_bn<32>(A, W, e, st);
}
