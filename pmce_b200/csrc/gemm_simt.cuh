// FP32 CUDA-core GEMM with fused epilogue:  out = epi( A[M,K] * W[N,K]^T ).
// Both operands are K-contiguous (nn.Linear weight layout [out,in] is used as is).
// This is the exact-fp32 path: single-pass TF32 tensor-core math misses the 1e-3 parity gate on the
// 40-GEMM-deep forward (measured in DESIGN.md §precision), so fp32 accumulate-and-multiply is the
// reference-accurate baseline every tensor-core kernel is validated against.
#pragma once
#include "common.cuh"

struct GemmEpi {
    const float* bias;     // [N] or nullptr
    const float* resid;    // addressed like `out`, or nullptr
    const float* rowadd;   // [rowadd_period, N] (ld = N), added at row % period; or nullptr
    int rowadd_period;
    int act;               // 0 none, 1 exact GELU
    int a_relu;            // apply relu to A on load
    RowMap rmap;           // output row -> offset
    RowMap cmap;           // output col -> offset
};

static inline GemmEpi gemm_epi_plain(int ldc) {
    GemmEpi e;
    e.bias = nullptr; e.resid = nullptr; e.rowadd = nullptr; e.rowadd_period = 1; e.act = 0; e.a_relu = 0;
    e.rmap.div = 1; e.rmap.s0 = ldc; e.rmap.s1 = 0;
    e.cmap.div = 1; e.cmap.s0 = 1; e.cmap.s1 = 0;
    return e;
}

// Thread tile TM x TN is split in 4-wide chunks strided across the block tile so that the per-k shared
// loads of a warp are contiguous float4s (no bank conflicts).
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, float* __restrict__ out,
               int M, int N, int K, GemmEpi e) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TX = BN / TN;           // threads along N
    constexpr int TY = BM / TM;           // threads along M
    constexpr int CM = TM / 4, CN = TN / 4;
    static_assert(TM % 4 == 0 && TN % 4 == 0 && BK % 4 == 0, "tile");
    constexpr int A_F4 = BM * BK / 4, W_F4 = BN * BK / 4;
    constexpr int A_PT = (A_F4 + NT - 1) / NT, W_PT = (W_F4 + NT - 1) / NT;

    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_PT], rw[W_PT];

    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PT; ++i) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < A_F4) {
                int row = idx / (BK / 4), kq = idx % (BK / 4);
                int gm = m0 + row, gk = k0 + kq * 4;
                if (gm < M && gk < K) v = ld4(A + (size_t)gm * lda + gk);
                if (e.a_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < W_PT; ++i) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < W_F4) {
                int row = idx / (BK / 4), kq = idx % (BK / 4);
                int gn = n0 + row, gk = k0 + kq * 4;
                if (gn < N && gk < K) v = ld4(W + (size_t)gn * ldw + gk);
            }
            rw[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PT; ++i) {
            int idx = tid + i * NT;
            if (idx < A_F4) {
                int row = idx / (BK / 4), kq = idx % (BK / 4);
                As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PT; ++i) {
            int idx = tid + i * NT;
            if (idx < W_F4) {
                int row = idx / (BK / 4), kq = idx % (BK / 4);
                Ws[buf][kq * 4 + 0][row] = rw[i].x; Ws[buf][kq * 4 + 1][row] = rw[i].y;
                Ws[buf][kq * 4 + 2][row] = rw[i].z; Ws[buf][kq * 4 + 3][row] = rw[i].w;
            }
        }
    };

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                float4 v = ld4(&As[buf][k][c * (TY * 4) + ty * 4]);
                a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                float4 v = ld4(&Ws[buf][k][c * (TX * 4) + tx * 4]);
                b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + (i / 4) * (TY * 4) + ty * 4 + (i % 4);
        if (gm >= M) continue;
        const long long roff = e.rmap(gm);
        const float* radd = e.rowadd ? e.rowadd + (size_t)(gm % e.rowadd_period) * N : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + (j / 4) * (TX * 4) + tx * 4 + (j % 4);
            if (gn >= N) continue;
            float v = acc[i][j];
            if (e.bias) v += e.bias[gn];
            if (e.act == 1) v = gelu_erf(v);
            if (radd) v += radd[gn];
            const long long off = roff + e.cmap(gn);
            if (e.resid) v += e.resid[off];
            out[off] = v;
        }
    }
}

// Host-side launcher: picks a tile shape that fills the 148 SMs.
static inline cudaError_t launch_gemm_tn(const float* A, int lda, const float* W, int ldw, float* out, int M, int N, int K,
                                         const GemmEpi& e, cudaStream_t st) {
    auto ctas = [&](int bm, int bn) { return (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn); };
    if (ctas(128, 128) >= 2 * 148) {
        dim3 g((N + 127) / 128, (M + 127) / 128);
        gemm_tn_kernel<128, 128, 16, 8, 8><<<g, 256, 0, st>>>(A, lda, W, ldw, out, M, N, K, e);
    } else if (ctas(64, 64) >= 148) {
        dim3 g((N + 63) / 64, (M + 63) / 64);
        gemm_tn_kernel<64, 64, 16, 4, 4><<<g, 256, 0, st>>>(A, lda, W, ldw, out, M, N, K, e);
    } else {
        dim3 g((N + 31) / 32, (M + 31) / 32);
        gemm_tn_kernel<32, 32, 32, 4, 4><<<g, 64, 0, st>>>(A, lda, W, ldw, out, M, N, K, e);
    }
    return cudaGetLastError();
}
