// tcgen05 / TMA / TMEM GEMM for sm_100a with split-bf16 ("bf16x3") operands and a fused epilogue:
//     out[M,N] = epi( A[M,K] * W[N,K]^T ),   A = A_hi + A_lo,  W = W_hi + W_lo  (bf16 pairs, K-contiguous)
// Persistent: one CTA per SM loops over 128 x BN output tiles (n fastest, so concurrently running CTAs share the A
// tile through L2). Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..9 = epilogue (TMEM lane group = warp % 4, column half = (warp-2)/4; one output row per thread).
// Three pipelines: smem ring of STAGES x {A_hi, A_lo, W_hi, W_lo} tiles ([rows][64 bf16], written by TMA with the
// 128-byte swizzle the UMMA descriptors expect) with full/empty mbarriers; TWO TMEM accumulators with
// tmem_full/tmem_empty mbarriers so the epilogue of tile i overlaps the MMAs of tile i+1; the tile loop itself.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"

struct TcEpi {
    const float* bias;      // [N] or null
    const float* resid;     // fp32 [M, ld_resid] or null
    const float* rowadd;    // [period, N] or null (added at row % period)
    int rowadd_period;
    int ld_resid;
    int act;                // 0 none, 1 exact GELU
    float* out_f32;         // fp32 [M, ld_out] or null
    __nv_bfloat16* out_hi;  // split output [M, ld_split] or null (both hi and lo, or neither)
    __nv_bfloat16* out_lo;
    int ld_out, ld_split;
    int mapped;             // out_f32 / resid are addressed as rmap(row) + cmap(col) instead of row*ld + col
    RowMap rmap, cmap;
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;            // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_THREADS = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;

template <int BN>
struct TcCfg {
    static constexpr int A_TILE = TC_BM * 128;          // bytes per A half (hi or lo)
    static constexpr int W_TILE = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * W_TILE;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 4 ? 4 : (200 * 1024) / STAGE_BYTES;
    static constexpr int STG_BYTES = TC_EPI_WARPS * 32 * 32 * 4;     // per-warp epilogue transpose tiles
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators
    static_assert(STAGES >= 2, "tile too large");
    static_assert(TMEM_COLS <= 512, "TMEM");
};

// Epilogue of one 32-row x 32-column accumulator chunk, warp-collective. The TMEM load gives each thread one ROW
// (32 consecutive columns); writing that straight to global memory touches 32 different 128-byte lines per instruction
// (measured: ~20 us per 128x256 tile, the GEMM's bottleneck). So the chunk is transposed through a 4 KB XOR-swizzled
// shared-memory tile: afterwards lane = column, and every global access (bias, row-embedding, residual read, fp32 and
// split-bf16 stores) is one fully coalesced row segment.
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpi& e, const uint32_t (&v)[32], uint32_t* stg, int lane, int row0, int col0,
                                                  int M, int N) {
#pragma unroll
    for (int i = 0; i < 32; ++i) stg[lane * 32 + (i ^ lane)] = v[i];
    __syncwarp();
    const int col = col0 + lane;
    const bool col_ok = col < N;
    const float b = (e.bias && col_ok) ? e.bias[col] : 0.f;
    const long long coff = e.mapped ? e.cmap(col) : (long long)col;
    const int nrows = M - row0 < 32 ? M - row0 : 32;
#pragma unroll 4
    for (int rr = 0; rr < nrows; ++rr) {
        float x = __uint_as_float(stg[rr * 32 + (lane ^ rr)]) + b;
        if (e.act == 1) x = gelu_erf(x);
        if (col_ok) {
            const int row = row0 + rr;
            if (e.rowadd) x += e.rowadd[(size_t)(row % e.rowadd_period) * N + col];
            if (e.mapped) {
                const long long off = e.rmap(row) + coff;
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
    __syncwarp();   // staging tile is reused by the next chunk
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 int M, int N, int K, TcEpi e) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t* stg_all = reinterpret_cast<uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (K + TC_BK - 1) / TC_BK;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + TC_BM - 1) / TC_BM;
    const int num_tiles = tiles_n * tiles_m;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA_hi); tc::tma_prefetch_desc(&tmA_lo);
        tc::tma_prefetch_desc(&tmW_hi); tc::tma_prefetch_desc(&tmW_lo);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], TC_EPI_WARPS); }
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
            uint32_t it = 0;                                    // global k-block counter across tiles
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    tc::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    tc::tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                    tc::tma_load_2d(st + Cfg::A_TILE, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                    tc::tma_load_2d(st + 2 * Cfg::A_TILE, &tmW_hi, &full_bar[s], kb * TC_BK, n0);
                    tc::tma_load_2d(st + 2 * Cfg::A_TILE + Cfg::W_TILE, &tmW_lo, &full_bar[s], kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(TC_BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                tc::mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
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
                        tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_lo, k), tc::umma_desc_advance_k(w_hi, k), idesc, (kb | k) != 0);
                        tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_lo, k), idesc, 1);
                        tc::umma_bf16(tmem_d, tc::umma_desc_advance_k(a_hi, k), tc::umma_desc_advance_k(w_hi, k), idesc, 1);
                    }
                    tc::umma_commit(&empty_bar[s]);          // smem slot free once these MMAs have read it
                }
                tc::umma_commit(&tmem_full_bar[acc]);        // accumulator complete
            }
        }
    } else {
        // ---- epilogue: 8 warps; lane group q = warp % 4 (TMEM lanes [32q,32q+32) = tile rows), column half = (warp-2)/4 ----
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CHUNKS = BN / 32;                       // 32-column chunks per tile
        constexpr int C_BEGIN_STRIDE = (CHUNKS + 1) / 2;      // chunks [0, C) for half 0, [C, CHUNKS) for half 1
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * BN;
            const uint32_t acc = tcount & 1;
            const int row0 = m0 + q * 32;
            uint32_t* stg = stg_all + (warp - 2) * 1024;
            tc::mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
            tc::tc_fence_after();
            const int cb = half == 0 ? 0 : C_BEGIN_STRIDE, ce = half == 0 ? C_BEGIN_STRIDE : CHUNKS;
#pragma unroll 1
            for (int c = cb; c < ce; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                tc::tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (row0 < M && col0 < N) tc_epilogue_chunk(e, v, stg, lane, row0, col0, M, N);   // warp-uniform condition
            }
            tc::tc_fence_before();
            if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);      // this warp is done reading the accumulator
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// fp32 [rows, cols] (ld) -> bf16 hi / lo [rows, ld_out]; optional relu on the way (linear_cur input).
__global__ void split_rows_kernel(const float* __restrict__ x, int rows, int cols, int ld, int relu, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ld_out) {
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

// bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows][64], 128-byte swizzle, zero OOB fill.
static inline int make_tmap_bf16(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

struct TcOperand {   // split bf16 matrix [rows, cols], row stride ld (elements)
    const __nv_bfloat16* hi; const __nv_bfloat16* lo; int rows, cols, ld;
};

static inline int tc_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN>
static inline int launch_linear_tc_bn(const TcOperand& A, const TcOperand& W, const TcEpi& e, cudaStream_t st) {
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    if (make_tmap_bf16(&ta_hi, A.hi, A.rows, A.cols, A.ld, TC_BM) || make_tmap_bf16(&ta_lo, A.lo, A.rows, A.cols, A.ld, TC_BM) ||
        make_tmap_bf16(&tw_hi, W.hi, W.rows, W.cols, W.ld, BN) || make_tmap_bf16(&tw_lo, W.lo, W.rows, W.cols, W.ld, BN))
        return 1;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(linear_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES) != cudaSuccess) return 2;
        configured = true;
    }
    const long long tiles = (long long)((W.rows + BN - 1) / BN) * ((A.rows + TC_BM - 1) / TC_BM);
    const int grid = (int)(tiles < tc_num_sms() ? tiles : tc_num_sms());
    linear_tc_kernel<BN><<<grid, TC_THREADS, TcCfg<BN>::SMEM_BYTES, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, A.rows, W.rows, A.cols, e);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// Tile width: the widest BN that still yields at least one wave of tiles; skinny problems take the narrowest tile so the
// weight stream is spread over as many SMs as possible.
// Requirements: A.cols == W.cols (K), K % 8 == 0, ld % 8 == 0, 16-byte aligned bases.
static inline int launch_linear_tc(const TcOperand& A, const TcOperand& W, const TcEpi& e, cudaStream_t st) {
    const int N = W.rows;
    const long long tiles_m = (A.rows + TC_BM - 1) / TC_BM;
    const int sms = tc_num_sms();
    static int forced = -1;   // PMCE_TC_BN=<32|64|128|256>: tile-sweep knob for profiling (tools/gemm_sweep.py)
    if (forced < 0) { const char* s = getenv("PMCE_TC_BN"); forced = s ? atoi(s) : 0; }
    if (forced == 256) return launch_linear_tc_bn<256>(A, W, e, st);
    if (forced == 128) return launch_linear_tc_bn<128>(A, W, e, st);
    if (forced == 64) return launch_linear_tc_bn<64>(A, W, e, st);
    if (forced == 32) return launch_linear_tc_bn<32>(A, W, e, st);
    if (N >= 256 && tiles_m * ((N + 255) / 256) >= sms) return launch_linear_tc_bn<256>(A, W, e, st);
    if (N >= 128 && tiles_m * ((N + 127) / 128) >= sms) return launch_linear_tc_bn<128>(A, W, e, st);
    if (N >= 64 && tiles_m * ((N + 63) / 64) >= sms) return launch_linear_tc_bn<64>(A, W, e, st);
    if (N > 32 && tiles_m * ((N + 63) / 64) * 2 > sms) return launch_linear_tc_bn<64>(A, W, e, st);
    return launch_linear_tc_bn<32>(A, W, e, st);
}
