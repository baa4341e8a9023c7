// Row-wise / attention / recurrence kernels of the PMCE hot path (fp32, CUDA cores, warp-shuffle reductions).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

// Split-bf16 activation pair: tensors that feed a tensor-core GEMM are written as hi/lo bf16.
struct SplitOut {
    __nv_bfloat16* hi; __nv_bfloat16* lo;
};
__device__ __forceinline__ void store_split4(const SplitOut& o, size_t idx, float4 v) {
    uint2 h, l;
    tc::split_bf16x2(v.x, v.y, h.x, l.x);
    tc::split_bf16x2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(o.hi + idx) = h;
    *reinterpret_cast<uint2*>(o.lo + idx) = l;
}
__device__ __forceinline__ void store_split1(const SplitOut& o, size_t idx, float a) {
    uint32_t h, l;
    tc::split_bf16x2(a, 0.f, h, l);          // first element sits in the low half
    *reinterpret_cast<unsigned short*>(o.hi + idx) = (unsigned short)(h & 0xffffu);
    *reinterpret_cast<unsigned short*>(o.lo + idx) = (unsigned short)(l & 0xffffu);
}
__device__ __forceinline__ void store_split2(const SplitOut& o, size_t idx, float a, float b) {
    uint32_t h, l;
    tc::split_bf16x2(a, b, h, l);
    *reinterpret_cast<uint32_t*>(o.hi + idx) = h;
    *reinterpret_cast<uint32_t*>(o.lo + idx) = l;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm over C features, one warp per row, C % 128 == 0, C <= 1024.
//   y1 = LN_a(x) (+ pos[(row / pos_div) % pos_mod])        -> out1 (optional)
//   y2 = LN_b(y1)                                           -> out2 (optional)
// Used for: norm_s / norm_t fused with the next block's norm1 (PoseEstimation.py:84-107), norm2.
// ------------------------------------------------------------------------------------------------------
struct LnParams {
    const float* w; const float* b; float eps;
};

template <int MAXV>  // MAXV = C / 128 upper bound
__device__ __forceinline__ void ln_inreg(float4 (&v)[MAXV], int nv, int C, const float* __restrict__ w,
                                         const float* __restrict__ b, float eps, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + bb * bb) + (c * c + d * d);
        }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int col = (i * 32 + lane) * 4;
            float4 ww = ld4(w + col), bb = ld4(b + col);
            v[i].x = (v[i].x - mean) * rstd * ww.x + bb.x;
            v[i].y = (v[i].y - mean) * rstd * ww.y + bb.y;
            v[i].z = (v[i].z - mean) * rstd * ww.z + bb.z;
            v[i].w = (v[i].w - mean) * rstd * ww.w + bb.w;
        }
}

template <int MAXV>   // C / 128 rounded up to a power of two (1, 2, 4, 8): the row lives in MAXV float4 per lane
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, int nrows, int C, LnParams a, int has_a, const float* __restrict__ pos,
               int pos_div, int pos_mod, float* __restrict__ out1, LnParams bparm, float* __restrict__ out2, SplitOut out2s,
               int map_rows, int map_stride) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int nv = C / 128;
    float4 v[MAXV];
    // map_rows > 0: output row r (window-major) is read from source row (r / map_rows) * map_stride + r % map_rows
    // (frame-major tokens of overlapping windows, pmce_forward_sliding)
    const size_t srow = map_rows > 0 ? (size_t)(row / map_rows) * map_stride + row % map_rows : (size_t)row;
    const float* xr = x + srow * C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) v[i] = ld4(xr + (i * 32 + lane) * 4);
    if (has_a) ln_inreg<MAXV>(v, nv, C, a.w, a.b, a.eps, lane);
    if (pos) {
        const float* pr = pos + (size_t)((row / pos_div) % pos_mod) * C;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                float4 p = ld4(pr + (i * 32 + lane) * 4);
                v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w;
            }
    }
    if (out1) {
        float* o = out1 + (size_t)row * C;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) st4(o + (i * 32 + lane) * 4, v[i]);
    }
    if (out2 || out2s.hi) {
        ln_inreg<MAXV>(v, nv, C, bparm.w, bparm.b, bparm.eps, lane);
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                const size_t idx = (size_t)row * C + (i * 32 + lane) * 4;
                if (out2) st4(out2 + idx, v[i]);
                if (out2s.hi) store_split4(out2s, idx, v[i]);
            }
    }
}

// ------------------------------------------------------------------------------------------------------
// Lifter embedding (PoseEstimation.py:78-81) fused with SpatialBlocks[0].norm1:
//   x0[b,t,j,:] = W_je p2d[b,t,j,:] + b_je + imgemb[b,t,:] + spos[j,:];  xn = LN(x0)
// imgemb already contains b_if (GEMM bias).  One warp per token.
// ------------------------------------------------------------------------------------------------------
template <int MAXV>   // as ln_rows_kernel: C / 128 rounded up to a power of two
__global__ void __launch_bounds__(256)
lifter_embed_kernel(const float* __restrict__ pose2d, const float* __restrict__ imgemb, const float* __restrict__ wje,
                    const float* __restrict__ bje, const float* __restrict__ spos, int ntok, int J, int C, LnParams n1,
                    float* __restrict__ x0, float* __restrict__ xn, SplitOut xns) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= ntok) return;
    const int nv = C / 128;
    const int j = row % J, bt = row / J;
    const float p0 = pose2d[(size_t)row * 2 + 0], p1 = pose2d[(size_t)row * 2 + 1];
    float4 v[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int col = (i * 32 + lane) * 4;
            float4 e = ld4(imgemb + (size_t)bt * C + col);
            float4 sp = ld4(spos + (size_t)j * C + col);
            float4 bb = ld4(bje + col);
            // wje is [C,2] row-major: 8 consecutive floats for 4 channels
            float4 w01 = ld4(wje + (size_t)col * 2), w23 = ld4(wje + (size_t)col * 2 + 4);
            float4 r;
            r.x = ((w01.x * p0 + w01.y * p1) + bb.x) + e.x + sp.x;
            r.y = ((w01.z * p0 + w01.w * p1) + bb.y) + e.y + sp.y;
            r.z = ((w23.x * p0 + w23.y * p1) + bb.z) + e.z + sp.z;
            r.w = ((w23.z * p0 + w23.w * p1) + bb.w) + e.w + sp.w;
            v[i] = r;
            st4(x0 + (size_t)row * C + col, r);
        }
    ln_inreg<MAXV>(v, nv, C, n1.w, n1.b, n1.eps, lane);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const size_t idx = (size_t)row * C + (i * 32 + lane) * 4;
            if (xn) st4(xn + idx, v[i]);
            if (xns.hi) store_split4(xns, idx, v[i]);
        }
}

// ------------------------------------------------------------------------------------------------------
// Lifter head (PoseEstimation.py:107-113): x = norm_t(y); r = W_r LN_1e-5(x) + b_r  (one warp per token),
// then fusion over frames + /1000 (PMCE.py:18) in lifter_fuse_kernel.
// ------------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
lifter_head_kernel(const float* __restrict__ y, int ntok, int C, LnParams nt, LnParams nh, const float* __restrict__ wr,
                   const float* __restrict__ br, float* __restrict__ r3) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= ntok) return;
    const int nv = C / 128;
    float4 v[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) v[i] = ld4(y + (size_t)row * C + (i * 32 + lane) * 4);
    ln_inreg<MAXV>(v, nv, C, nt.w, nt.b, nt.eps, lane);
    ln_inreg<MAXV>(v, nv, C, nh.w, nh.b, nh.eps, lane);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int col = (i * 32 + lane) * 4;
            float4 w0 = ld4(wr + col), w1 = ld4(wr + C + col), w2 = ld4(wr + 2 * C + col);
            a0 += v[i].x * w0.x + v[i].y * w0.y + v[i].z * w0.z + v[i].w * w0.w;
            a1 += v[i].x * w1.x + v[i].y * w1.y + v[i].z * w1.z + v[i].w * w1.w;
            a2 += v[i].x * w2.x + v[i].y * w2.y + v[i].z * w2.z + v[i].w * w2.w;
        }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        r3[(size_t)row * 3 + 0] = a0 + br[0];
        r3[(size_t)row * 3 + 1] = a1 + br[1];
        r3[(size_t)row * 3 + 2] = a2 + br[2];
    }
}

// pose3d[b,j,c] = sum_t w[t] r[b,t,j,c] + bias ; joints = pose3d / 1000
__global__ void lifter_fuse_kernel(const float* __restrict__ r3, const float* __restrict__ wf, const float* __restrict__ bf,
                                   int B, int T, int J, float* __restrict__ pose3d, float* __restrict__ joints_m) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = B * J * 3;
    if (idx >= n) return;
    const int c = idx % 3, j = (idx / 3) % J, b = idx / (3 * J);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(wf[t], r3[((size_t)(b * T + t) * J + j) * 3 + c], acc);
    acc += bf[0];
    pose3d[idx] = acc;
    if (joints_m) joints_m[idx] = __fdiv_rn(acc, 1000.0f);
}

// ------------------------------------------------------------------------------------------------------
// Generic multi-head attention core: O = softmax(Q K^T * scale) V per (sequence, head).
// One CTA = (query tile of 32*NW, head, sequence); the head's K and V slices live in shared memory and are
// read as warp broadcasts; each lane owns one query row and runs an online softmax over the keys.
// Row addressing is generalised so the same kernel serves the lifter's spatial / temporal passes
// (strided sequences inside [B,T,J,*]) and the decoder's cross/self attention.
// ------------------------------------------------------------------------------------------------------
struct AttnAddr {
    RowMap seq;        // sequence index -> first row
    long long tok;     // token stride in rows
    int ld;            // row stride in floats
};

template <int D>
__global__ void __launch_bounds__(128)
attn_kernel(const float* __restrict__ Q, AttnAddr aq, const float* __restrict__ K, const float* __restrict__ V, AttnAddr akv,
            float* __restrict__ O, SplitOut Os, AttnAddr ao, int N1, int N2, float scale) {
    pdl_enter();
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                       // [N2][D]
    float* Vs = smem + (size_t)N2 * D;      // [N2][D]
    const int h = blockIdx.y, s = blockIdx.z;
    const int tid = threadIdx.x;

    const long long kv0 = akv.seq(s);
    constexpr int D4 = D / 4;
    for (int idx = tid; idx < N2 * D4; idx += blockDim.x) {
        const int r = idx / D4, c = (idx % D4) * 4;
        const size_t g = (size_t)(kv0 + (long long)r * akv.tok) * akv.ld + h * D + c;
        st4(Ks + r * D + c, ld4(K + g));
        st4(Vs + r * D + c, ld4(V + g));
    }
    __syncthreads();

    const int qi = blockIdx.x * blockDim.x + tid;
    if (qi >= N1) return;
    const float* qp = Q + (size_t)(aq.seq(s) + (long long)qi * aq.tok) * aq.ld + h * D;
    float q[D], o[D];
#pragma unroll
    for (int c = 0; c < D; c += 4) {
        float4 v = ld4(qp + c);
        q[c] = v.x * scale; q[c + 1] = v.y * scale; q[c + 2] = v.z * scale; q[c + 3] = v.w * scale;
        o[c] = o[c + 1] = o[c + 2] = o[c + 3] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    int k = 0;
    // keys in blocks of 4: one running-max update / accumulator rescale per block instead of per key
    for (; k + 4 <= N2; k += 4) {
        float sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* kr = Ks + (k + j) * D;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                float4 kk = ld4(kr + c);
                s0 = fmaf(q[c], kk.x, s0); s1 = fmaf(q[c + 1], kk.y, s1);
                s0 = fmaf(q[c + 2], kk.z, s0); s1 = fmaf(q[c + 3], kk.w, s1);
            }
            sc[j] = s0 + s1;
        }
        const float mn = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
        const float alpha = expf(m - mn);
        m = mn;
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = expf(sc[j] - mn);
        l = l * alpha + ((p[0] + p[1]) + (p[2] + p[3]));
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] *= alpha;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* vr = Vs + (k + j) * D;
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                float4 vv = ld4(vr + c);
                o[c] = fmaf(p[j], vv.x, o[c]); o[c + 1] = fmaf(p[j], vv.y, o[c + 1]);
                o[c + 2] = fmaf(p[j], vv.z, o[c + 2]); o[c + 3] = fmaf(p[j], vv.w, o[c + 3]);
            }
        }
    }
    for (; k < N2; ++k) {
        const float* kr = Ks + k * D;
        float sc0 = 0.f, sc1 = 0.f;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
            float4 kk = ld4(kr + c);
            sc0 = fmaf(q[c], kk.x, sc0); sc1 = fmaf(q[c + 1], kk.y, sc1);
            sc0 = fmaf(q[c + 2], kk.z, sc0); sc1 = fmaf(q[c + 3], kk.w, sc1);
        }
        const float sc = sc0 + sc1;
        const float mn = fmaxf(m, sc);
        const float alpha = expf(m - mn);
        const float p = expf(sc - mn);
        l = l * alpha + p;
        m = mn;
        const float* vr = Vs + k * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
            float4 vv = ld4(vr + c);
            o[c] = fmaf(o[c], alpha, p * vv.x); o[c + 1] = fmaf(o[c + 1], alpha, p * vv.y);
            o[c + 2] = fmaf(o[c + 2], alpha, p * vv.z); o[c + 3] = fmaf(o[c + 3], alpha, p * vv.w);
        }
    }
    const float inv = 1.0f / l;
    const size_t obase = (size_t)(ao.seq(s) + (long long)qi * ao.tok) * ao.ld + h * D;
#pragma unroll
    for (int c = 0; c < D; c += 4) {
        const float4 r = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
        if (O) st4(O + obase + c, r);
        if (Os.hi) store_split4(Os, obase + c, r);
    }
}

// ------------------------------------------------------------------------------------------------------
// Few queries, many keys: the joint stream's cross-attention (CoevoDecoder.py:167 joint_CA_FFN -> :47-62: 17 queries x 431 keys,
// 8 heads of 8). attn_kernel gives every query ONE lane that walks all keys (17 lanes x 431 dependent softmax updates = 37 us at
// 64 clips); here a CTA owns (head, sequence), a warp owns a query and the KEYS are spread over its lanes: each lane keeps the
// scores of its <= KPL keys in registers (exact two-pass softmax), the row max / sum / output are warp reductions.
// Shared rows are padded to D + 4 floats so a quarter-warp's 16-byte reads fall into distinct banks.
// ------------------------------------------------------------------------------------------------------
constexpr int ATTN_FEWQ_KPL = 16;        // keys per lane: N2 <= 512
template <int D>
__global__ void __launch_bounds__(256)
attn_fewq_kernel(const float* __restrict__ Q, AttnAddr aq, const float* __restrict__ K, const float* __restrict__ V, AttnAddr akv,
                 float* __restrict__ O, SplitOut Os, AttnAddr ao, int N1, int N2, float scale) {
    pdl_enter();
    constexpr int LD = D + 4, D4 = D / 4;
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                        // [N2][LD]
    float* Vs = smem + (size_t)N2 * LD;      // [N2][LD]
    const int h = blockIdx.x, s = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const long long kv0 = akv.seq(s);
    for (int idx = tid; idx < N2 * D4; idx += blockDim.x) {
        const int r = idx / D4, c = (idx % D4) * 4;
        const size_t g = (size_t)(kv0 + (long long)r * akv.tok) * akv.ld + h * D + c;
        st4(Ks + r * LD + c, ld4(K + g));
        st4(Vs + r * LD + c, ld4(V + g));
    }
    __syncthreads();
    for (int qi = warp; qi < N1; qi += nw) {
        const float* qp = Q + (size_t)(aq.seq(s) + (long long)qi * aq.tok) * aq.ld + h * D;
        float q[D];
#pragma unroll
        for (int c = 0; c < D; c += 4) {
            const float4 v = ld4(qp + c);
            q[c] = v.x * scale; q[c + 1] = v.y * scale; q[c + 2] = v.z * scale; q[c + 3] = v.w * scale;
        }
        float sc[ATTN_FEWQ_KPL];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < ATTN_FEWQ_KPL; ++j) {
            const int k = lane + 32 * j;
            sc[j] = -INFINITY;
            if (k < N2) {
                const float* kr = Ks + k * LD;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < D; c += 4) {
                    const float4 kk = ld4(kr + c);
                    s0 = fmaf(q[c], kk.x, s0); s1 = fmaf(q[c + 1], kk.y, s1);
                    s0 = fmaf(q[c + 2], kk.z, s0); s1 = fmaf(q[c + 3], kk.w, s1);
                }
                sc[j] = s0 + s1;
            }
            m = fmaxf(m, sc[j]);
        }
        m = warp_max(m);
        float l = 0.f, o[D];
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] = 0.f;
#pragma unroll
        for (int j = 0; j < ATTN_FEWQ_KPL; ++j) {
            const int k = lane + 32 * j;
            if (k < N2) {
                const float pj = expf(sc[j] - m);
                l += pj;
                const float* vr = Vs + k * LD;
#pragma unroll
                for (int c = 0; c < D; c += 4) {
                    const float4 vv = ld4(vr + c);
                    o[c] = fmaf(pj, vv.x, o[c]); o[c + 1] = fmaf(pj, vv.y, o[c + 1]);
                    o[c + 2] = fmaf(pj, vv.z, o[c + 2]); o[c + 3] = fmaf(pj, vv.w, o[c + 3]);
                }
            }
        }
        l = warp_sum(l);
        float mine = 0.f;                    // lane c keeps output element c (c, c + 32, ... for D > 32)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float t = warp_sum(o[c]);
            if ((c & 31) == lane) {
                const size_t oi = (size_t)(ao.seq(s) + (long long)qi * ao.tok) * ao.ld + h * D + c;
                mine = t / l;
                if (O) O[oi] = mine;
                if (Os.hi) {
                    __nv_bfloat16 hi, lo;
                    tc::split_bf16(mine, hi, lo);
                    Os.hi[oi] = hi; Os.lo[oi] = lo;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// AdaLayerNorm apply (CoevoDecoder.py:23-29), 64 features per row, one warp per row:
//   y = gamma_b * (x - mean) / (std_unbiased + eps) + beta_b,  gamma/beta = gb[b, slot, 0:64 | 64:128]
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adaln_apply_kernel(const float* __restrict__ x, int nrows, int rows_per_batch, const float* __restrict__ gb, int gb_ld,
                   int slot, float eps, float* __restrict__ y, SplitOut ys) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int b = row / rows_per_batch;
    const float2 v = *reinterpret_cast<const float2*>(x + (size_t)row * 64 + lane * 2);
    const float mean = warp_sum(v.x + v.y) * (1.0f / 64.0f);
    const float dx = v.x - mean, dy = v.y - mean;
    const float var = warp_sum(dx * dx + dy * dy) * (1.0f / 63.0f);
    const float inv = 1.0f / (sqrtf(var) + eps);
    const float* g = gb + (size_t)b * gb_ld + slot * 128;
    const float2 ga = *reinterpret_cast<const float2*>(g + lane * 2);
    const float2 be = *reinterpret_cast<const float2*>(g + 64 + lane * 2);
    float2 r;
    r.x = ga.x * dx * inv + be.x;
    r.y = ga.y * dy * inv + be.y;
    if (y) *reinterpret_cast<float2*>(y + (size_t)row * 64 + lane * 2) = r;
    if (ys.hi) store_split2(ys, (size_t)row * 64 + lane * 2, r.x, r.y);
}

// ------------------------------------------------------------------------------------------------------
// Coordinate -> feature embedding (CoevoDecoder.py:177-180): f = W[64,3] p + b + pos[i];
// out_f = f (optional), out_q = f + qemb[i] (optional). One thread per (row, 4 features).
// ------------------------------------------------------------------------------------------------------
__global__ void coevo_embed_kernel(const float* __restrict__ coords, int nrows, int ntok, const float* __restrict__ w,
                                   const float* __restrict__ bias, const float* __restrict__ pos,
                                   const float* __restrict__ qemb, float* __restrict__ out_f, SplitOut out_fs, float* __restrict__ out_q) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * 16) return;
    const int row = idx >> 4, c = (idx & 15) * 4;
    const int i = row % ntok;
    const float p0 = coords[(size_t)row * 3], p1 = coords[(size_t)row * 3 + 1], p2 = coords[(size_t)row * 3 + 2];
    float f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float* wr = w + (c + u) * 3;
        f[u] = ((wr[0] * p0 + wr[1] * p1) + wr[2] * p2) + bias[c + u] + pos[(size_t)i * 64 + c + u];
    }
    if (out_f) st4(out_f + (size_t)row * 64 + c, make_float4(f[0], f[1], f[2], f[3]));
    if (out_fs.hi) store_split4(out_fs, (size_t)row * 64 + c, make_float4(f[0], f[1], f[2], f[3]));
    if (out_q) {
        float4 qe = ld4(qemb + (size_t)i * 64 + c);
        st4(out_q + (size_t)row * 64 + c, make_float4(f[0] + qe.x, f[1] + qe.y, f[2] + qe.z, f[3] + qe.w));
    }
}

// feature -> coordinate projection + residual (CoevoDecoder.py:189): out = W[3,64] x + b + coords. Warp per row.
__global__ void __launch_bounds__(256)
feat2coor_kernel(const float* __restrict__ x, int nrows, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ coords, float* __restrict__ out) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const float2 v = *reinterpret_cast<const float2*>(x + (size_t)row * 64 + lane * 2);
    float a[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float2 ww = *reinterpret_cast<const float2*>(w + c * 64 + lane * 2);
        a[c] = warp_sum(v.x * ww.x + v.y * ww.y);
    }
    if (lane < 3) out[(size_t)row * 3 + lane] = (lane == 0 ? a[0] : (lane == 1 ? a[1] : a[2])) + bias[lane] + coords[(size_t)row * 3 + lane];
}

// verts0[b,i,:] = joints[b, vj[i], :]  (CoevoDecoder.py:232) — pure copy, bit exact.
__global__ void gather_verts_kernel(const float* __restrict__ joints, const int32_t* __restrict__ vj, int B, int J, int Vd,
                                    float* __restrict__ verts) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * Vd * 3) return;
    const int c = idx % 3, i = (idx / 3) % Vd, b = idx / (3 * Vd);
    verts[idx] = joints[((size_t)b * J + vj[i]) * 3 + c];
}

// im2col for upsample_conv (Conv1d(431->6890,k=3,pad=1) over the xyz axis, CoevoDecoder.py:214,238):
//   A[(b,l), c*3+k] = verts[b,c,l+k-1] (0 outside [0,3)), padded to ldk columns with zeros.
__global__ void upsample_im2col_kernel(const float* __restrict__ verts, int B, int Vd, int ldk, float* __restrict__ A, SplitOut As) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = B * 3 * ldk;
    if (idx >= total) return;
    const int kk = idx % ldk, l = (idx / ldk) % 3, b = idx / (3 * ldk);
    float v = 0.f;
    if (kk < Vd * 3) {
        const int c = kk / 3, k = kk % 3;
        const int src = l + k - 1;
        if (src >= 0 && src < 3) v = verts[((size_t)b * Vd + c) * 3 + src];
    }
    if (A) A[idx] = v;
    if (As.hi) { __nv_bfloat16 h, l; tc::split_bf16(v, h, l); As.hi[idx] = h; As.lo[idx] = l; }
}

// ------------------------------------------------------------------------------------------------------
// One GRU time step for up to two directions (PyTorch nn.GRU gate order r,z,n):
//   gh = W_hh h + b_hh ; r = s(gi_r+gh_r) ; z = s(gi_z+gh_z) ; n = tanh(gi_n + r*gh_n) ; h' = (1-z) n + z h
// CTA tile: 64 batch rows x 16 hidden units x 3 gates, K = H in chunks of 16; gate math fused in the epilogue.
// ------------------------------------------------------------------------------------------------------
struct GruDir {
    const float* gi;     // [B, 3H] rows for this step (ld_gi)
    const float* hprev;  // [B, H] (ld_h) or nullptr for h = 0
    const float* whh;    // [3H, H]
    const float* bhh;    // [3H]
    float* hout;         // [B, H] (ld_o)
    SplitOut hs;         // optional split copy of h' (ld_s) for the tensor-core consumers, or {nullptr,nullptr}
    int ld_gi, ld_h, ld_o, ld_s;
};

__global__ void __launch_bounds__(256)
gru_step_kernel(GruDir d0, GruDir d1, int B, int H) {
    pdl_enter();
    const GruDir d = blockIdx.z == 0 ? d0 : d1;
    constexpr int BM = 64, BJ = 16, BK = 16;
    __shared__ __align__(16) float Hs[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][3 * BJ + 4];
    const int tid = threadIdx.x;
    const int u = tid % BJ, rg = tid / BJ;      // 16 units x 16 row-groups of 4 rows
    const int j0 = blockIdx.x * BJ, m0 = blockIdx.y * BM;
    float acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;

    if (d.hprev) {
        for (int k0 = 0; k0 < H; k0 += BK) {
            {   // h tile: 64 rows x 16 k = 256 float4 -> one per thread
                const int row = tid / 4, kq = tid % 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m0 + row < B) v = ld4(d.hprev + (size_t)(m0 + row) * d.ld_h + k0 + kq * 4);
                Hs[kq * 4 + 0][row] = v.x; Hs[kq * 4 + 1][row] = v.y; Hs[kq * 4 + 2][row] = v.z; Hs[kq * 4 + 3][row] = v.w;
            }
            if (tid < 192) {  // W tile: 48 rows x 16 k = 192 float4
                const int wr = tid / 4, kq = tid % 4;
                const int gate = wr / BJ, uu = wr % BJ;
                float4 v = ld4(d.whh + (size_t)(gate * H + j0 + uu) * H + k0 + kq * 4);
                Ws[kq * 4 + 0][wr] = v.x; Ws[kq * 4 + 1][wr] = v.y; Ws[kq * 4 + 2][wr] = v.z; Ws[kq * 4 + 3][wr] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 hv = ld4(&Hs[k][rg * 4]);
                const float w0 = Ws[k][u], w1 = Ws[k][BJ + u], w2 = Ws[k][2 * BJ + u];
                acc[0][0] = fmaf(hv.x, w0, acc[0][0]); acc[0][1] = fmaf(hv.x, w1, acc[0][1]); acc[0][2] = fmaf(hv.x, w2, acc[0][2]);
                acc[1][0] = fmaf(hv.y, w0, acc[1][0]); acc[1][1] = fmaf(hv.y, w1, acc[1][1]); acc[1][2] = fmaf(hv.y, w2, acc[1][2]);
                acc[2][0] = fmaf(hv.z, w0, acc[2][0]); acc[2][1] = fmaf(hv.z, w1, acc[2][1]); acc[2][2] = fmaf(hv.z, w2, acc[2][2]);
                acc[3][0] = fmaf(hv.w, w0, acc[3][0]); acc[3][1] = fmaf(hv.w, w1, acc[3][1]); acc[3][2] = fmaf(hv.w, w2, acc[3][2]);
            }
            __syncthreads();
        }
    }
    const int j = j0 + u;
    const float br = d.bhh[j], bz = d.bhh[H + j], bn = d.bhh[2 * H + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + rg * 4 + i;
        if (row >= B) continue;
        const float* gi = d.gi + (size_t)row * d.ld_gi;
        const float r = sigmoid_f(gi[j] + (acc[i][0] + br));
        const float z = sigmoid_f(gi[H + j] + (acc[i][1] + bz));
        const float n = tanhf(gi[2 * H + j] + r * (acc[i][2] + bn));
        const float hp = d.hprev ? d.hprev[(size_t)row * d.ld_h + j] : 0.f;
        const float hn = (1.0f - z) * n + z * hp;
        d.hout[(size_t)row * d.ld_o + j] = hn;
        if (d.hs.hi) { __nv_bfloat16 hh, ll; tc::split_bf16(hn, hh, ll); d.hs.hi[(size_t)row * d.ld_s + j] = hh; d.hs.lo[(size_t)row * d.ld_s + j] = ll; }
    }
}

// ------------------------------------------------------------------------------------------------------
// Sparse J-regressor (core/base.py:225): out[b,r,c] = scale * sum_i vals[i] * mesh[b, cols[i], c]
// ------------------------------------------------------------------------------------------------------
__global__ void jregress_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                const float* __restrict__ vals, int R, const float* __restrict__ mesh, int V, int B,
                                float scale, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * R * 3) return;
    const int c = idx % 3, r = (idx / 3) % R, b = idx / (3 * R);
    float acc = 0.f;
    for (int i = row_ptr[r]; i < row_ptr[r + 1]; ++i) acc = fmaf(vals[i], mesh[((size_t)b * V + cols[i]) * 3 + c] * scale, acc);
    out[idx] = acc;
}

// ------------------------------------------------------------------------------------------------------
// Evaluation epilogue of the test loop (lib/core/base.py:223-227 + data/PW3D/dataset.py:269-282 compute_both_err), one CTA
// per clip: root-align predicted / target mesh and joints on joint 0, mean per-vertex L2 and mean per-evaluation-joint L2.
//   pred mesh = cam_mesh * scale, target mesh = gt_mesh * scale (metres -> mm), pred_pose (already in mm, from
//   jregress_kernel), gt_pose in mm.   clip_err[b] = (joint mean error, mesh mean error) of clip b.
// The reference does this with four device->host copies of [B,6890,3] tensors and numpy every batch.
// ------------------------------------------------------------------------------------------------------
constexpr int EVAL_THREADS = 512;
__global__ void __launch_bounds__(EVAL_THREADS)
eval_err_kernel(const float* __restrict__ cam_mesh, const float* __restrict__ gt_mesh, const float* __restrict__ pred_pose,
                const float* __restrict__ gt_pose, const int32_t* __restrict__ eval_joints, int n_eval, int R, int V, float scale,
                float* __restrict__ clip_err) {
    __shared__ float red[2][EVAL_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* pp = pred_pose + (size_t)b * R * 3;
    const float* gp = gt_pose + (size_t)b * R * 3;
    const float rpx = pp[0], rpy = pp[1], rpz = pp[2], rgx = gp[0], rgy = gp[1], rgz = gp[2];
    // root-aligned difference (p*s - rp) - (g*s - rg) evaluated as in the reference, two vertices (six floats = three
    // 8-byte loads per mesh; a clip's mesh starts 8-byte aligned when V is even) per thread and iteration
    const float2* pm = reinterpret_cast<const float2*>(cam_mesh + (size_t)b * V * 3);
    const float2* gm = reinterpret_cast<const float2*>(gt_mesh + (size_t)b * V * 3);
    float macc = 0.f;
    const int npair = V >> 1;
    for (int i = tid; i < npair; i += EVAL_THREADS) {
        const float2 p0 = pm[3 * i], p1 = pm[3 * i + 1], p2 = pm[3 * i + 2];
        const float2 g0 = gm[3 * i], g1 = gm[3 * i + 1], g2 = gm[3 * i + 2];
        const float ax = (p0.x * scale - rpx) - (g0.x * scale - rgx), ay = (p0.y * scale - rpy) - (g0.y * scale - rgy),
                    az = (p1.x * scale - rpz) - (g1.x * scale - rgz);
        const float bx = (p1.y * scale - rpx) - (g1.y * scale - rgx), by = (p2.x * scale - rpy) - (g2.x * scale - rgy),
                    bz = (p2.y * scale - rpz) - (g2.y * scale - rgz);
        macc += sqrtf((ax * ax + ay * ay) + az * az) + sqrtf((bx * bx + by * by) + bz * bz);
    }
    if ((V & 1) && tid == 0) {                      // odd vertex count: the last vertex, scalar loads
        const float* p = cam_mesh + ((size_t)b * V + V - 1) * 3;
        const float* g = gt_mesh + ((size_t)b * V + V - 1) * 3;
        const float dx = (p[0] * scale - rpx) - (g[0] * scale - rgx), dy = (p[1] * scale - rpy) - (g[1] * scale - rgy),
                    dz = (p[2] * scale - rpz) - (g[2] * scale - rgz);
        macc += sqrtf((dx * dx + dy * dy) + dz * dz);
    }
    float jacc = 0.f;
    if (tid < n_eval) {
        const int j = eval_joints[tid];
        const float dx = (pp[j * 3] - rpx) - (gp[j * 3] - rgx), dy = (pp[j * 3 + 1] - rpy) - (gp[j * 3 + 1] - rgy),
                    dz = (pp[j * 3 + 2] - rpz) - (gp[j * 3 + 2] - rgz);
        jacc = sqrtf((dx * dx + dy * dy) + dz * dz);
    }
    macc = warp_sum(macc); jacc = warp_sum(jacc);
    if ((tid & 31) == 0) { red[0][tid >> 5] = jacc; red[1][tid >> 5] = macc; }
    __syncthreads();
    if (tid < 2) {
        float t = 0.f;
        for (int w = 0; w < EVAL_THREADS / 32; ++w) t += red[tid][w];
        clip_err[(size_t)b * 2 + tid] = t / (float)(tid == 0 ? n_eval : V);
    }
}

// mean over clips (fixed order: deterministic) -> mean_err = (joint_mean_error, mesh_mean_error)
__global__ void __launch_bounds__(64)
eval_mean_kernel(const float* __restrict__ clip_err, int B, float* __restrict__ mean_err) {
    const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int b = lane; b < B; b += 32) acc += clip_err[(size_t)b * 2 + which];
    acc = warp_sum(acc);
    if (lane == 0) mean_err[which] = acc / (float)B;
}

// ------------------------------------------------------------------------------------------------------
// SMPL LBS (smpl_layer.py:65-158)
// kernel A, one thread block (32 threads) per sample: rodrigues, joint regression from betas, FK chain,
// rest-pose removal -> A[b,24,12] (3x4 per joint), coef[b, :] = [betas(10) | vec(R_1..23 - I)(207) | 0 pad], joints.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
smpl_pose_kernel(const float* __restrict__ pose, const float* __restrict__ betas, const float* __restrict__ trans,
                 const float* __restrict__ j_template, const float* __restrict__ j_shapedirs,
                 const int32_t* __restrict__ parents, int B, int ldc, float* __restrict__ coef, float* __restrict__ Aout,
                 float* __restrict__ joints, float out_scale, SplitOut coef_s) {
    const int b = blockIdx.x, lane = threadIdx.x;
    __shared__ float R[24][9];
    __shared__ float Jr[24][3];
    __shared__ float G[24][12];
    float* cf = coef + (size_t)b * ldc;
    if (lane < 10) cf[lane] = betas[(size_t)b * 10 + lane];
    if (lane < ldc - 217) cf[217 + lane] = 0.f;
    if (coef_s.hi) {      // the split-bf16 copy the tensor-core blend-shape GEMM reads (saves a split_rows launch)
        if (lane < 10) store_split1(coef_s, (size_t)b * ldc + lane, betas[(size_t)b * 10 + lane]);
        if (lane < ldc - 217) store_split1(coef_s, (size_t)b * ldc + 217 + lane, 0.f);
    }
    if (lane < 24) {
        // batch_rodrigues (rodrigues_layer.py:41-52): theta = |a + 1e-8|, axis = a / theta, quaternion -> matrix
        const float ax = pose[(size_t)b * 72 + lane * 3], ay = pose[(size_t)b * 72 + lane * 3 + 1], az = pose[(size_t)b * 72 + lane * 3 + 2];
        const float ex = ax + 1e-8f, ey = ay + 1e-8f, ez = az + 1e-8f;
        const float theta = sqrtf(ex * ex + ey * ey + ez * ez);
        const float nx = ax / theta, ny = ay / theta, nz = az / theta;
        const float half = theta * 0.5f;
        const float cs = cosf(half), sn = sinf(half);
        float w = cs, x = sn * nx, y = sn * ny, z = sn * nz;
        const float qn = sqrtf(w * w + x * x + y * y + z * z);
        w /= qn; x /= qn; y /= qn; z /= qn;
        const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
        const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
        R[lane][0] = w2 + x2 - y2 - z2; R[lane][1] = 2 * xy - 2 * wz;     R[lane][2] = 2 * wy + 2 * xz;
        R[lane][3] = 2 * wz + 2 * xy;     R[lane][4] = w2 - x2 + y2 - z2; R[lane][5] = 2 * yz - 2 * wx;
        R[lane][6] = 2 * xz - 2 * wy;     R[lane][7] = 2 * wx + 2 * yz;     R[lane][8] = w2 - x2 - y2 + z2;
        // rest joints: Jr = j_template + j_shapedirs . betas
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = j_template[lane * 3 + c];
            for (int k = 0; k < 10; ++k) a = fmaf(j_shapedirs[(lane * 3 + c) * 10 + k], betas[(size_t)b * 10 + k], a);
            Jr[lane][c] = a;
        }
        if (lane >= 1) {
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float v = R[lane][e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
                cf[10 + (lane - 1) * 9 + e] = v;
                if (coef_s.hi) store_split1(coef_s, (size_t)b * ldc + 10 + (lane - 1) * 9 + e, v);
            }
        }
    }
    __syncwarp();
    {
        // forward kinematics (smpl_layer.py:105-119); G = [R | t] 3x4, row-major. Element (r, c) of a joint depends only on row r
        // of its parent: lanes 0..11 own one element each and walk the chain together (same expression order as a serial loop)
        const int r = lane >> 2, c = lane & 3;
        for (int i = 0; i < 24; ++i) {
            if (lane < 12) {
                if (i == 0) {
                    G[0][lane] = c < 3 ? R[0][r * 3 + c] : Jr[0][r];
                } else {
                    const int p = parents[i];
                    const float g0 = G[p][r * 4], g1 = G[p][r * 4 + 1], g2 = G[p][r * 4 + 2], g3 = G[p][r * 4 + 3];
                    if (c < 3) G[i][lane] = (g0 * R[i][c] + g1 * R[i][3 + c]) + g2 * R[i][6 + c];
                    else G[i][lane] = ((g0 * (Jr[i][0] - Jr[p][0]) + g1 * (Jr[i][1] - Jr[p][1])) + g2 * (Jr[i][2] - Jr[p][2])) + g3;
                }
            }
            __syncwarp();
        }
    }
    if (lane < 24) {
        const float tx = trans ? trans[(size_t)b * 3] : 0.f, ty = trans ? trans[(size_t)b * 3 + 1] : 0.f, tz = trans ? trans[(size_t)b * 3 + 2] : 0.f;
        joints[((size_t)b * 24 + lane) * 3 + 0] = (G[lane][3] + tx) * out_scale;
        joints[((size_t)b * 24 + lane) * 3 + 1] = (G[lane][7] + ty) * out_scale;
        joints[((size_t)b * 24 + lane) * 3 + 2] = (G[lane][11] + tz) * out_scale;
        // A = G - [0 | G . (Jr,0)]  (smpl_layer.py:126-132)
        float* a = Aout + ((size_t)b * 24 + lane) * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float corr = (G[lane][r * 4] * Jr[lane][0] + G[lane][r * 4 + 1] * Jr[lane][1]) + G[lane][r * 4 + 2] * Jr[lane][2];
            a[r * 4 + 0] = G[lane][r * 4 + 0]; a[r * 4 + 1] = G[lane][r * 4 + 1]; a[r * 4 + 2] = G[lane][r * 4 + 2];
            a[r * 4 + 3] = G[lane][r * 4 + 3] - corr;
        }
    }
}

// kernel B, skinning (smpl_layer.py:134-155): T_v = sum_i w[v,i] A_i ; verts = T_v [v_posed;1] + trans.
// One thread per vertex, SMPL_SKIN_NB samples per CTA: the vertex's 24 skinning weights are read ONCE into registers and
// reused for every sample of the CTA (one sample per CTA re-read the 661 KB weight table per sample: 169 MB of L2 traffic at
// B=256 and 50 us; the transforms of the CTA's samples sit in shared memory and are read as 16-byte broadcasts).
constexpr int SMPL_SKIN_NB = 8;
__global__ void __launch_bounds__(256)
smpl_skin_kernel(const float* __restrict__ v_posed, const float* __restrict__ Amat, const float* __restrict__ weights,
                 const float* __restrict__ trans, int V, int B, int ld_vp, float* __restrict__ verts, float out_scale) {
    __shared__ __align__(16) float As[SMPL_SKIN_NB][24 * 12];
    const int b0 = blockIdx.y * SMPL_SKIN_NB;
    const int nb = B - b0 < SMPL_SKIN_NB ? B - b0 : SMPL_SKIN_NB;
    for (int i = threadIdx.x; i < nb * 288; i += blockDim.x) As[i / 288][i % 288] = Amat[(size_t)b0 * 288 + i];
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float w[24];
#pragma unroll
    for (int i = 0; i < 24; i += 4) {
        const float4 w4 = ld4(weights + (size_t)v * 24 + i);
        w[i] = w4.x; w[i + 1] = w4.y; w[i + 2] = w4.z; w[i + 3] = w4.w;
    }
    for (int s = 0; s < nb; ++s) {
        const int b = b0 + s;
        float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            const float4 a0 = ld4(&As[s][i * 12]), a1 = ld4(&As[s][i * 12 + 4]), a2 = ld4(&As[s][i * 12 + 8]);
            T0.x = fmaf(w[i], a0.x, T0.x); T0.y = fmaf(w[i], a0.y, T0.y); T0.z = fmaf(w[i], a0.z, T0.z); T0.w = fmaf(w[i], a0.w, T0.w);
            T1.x = fmaf(w[i], a1.x, T1.x); T1.y = fmaf(w[i], a1.y, T1.y); T1.z = fmaf(w[i], a1.z, T1.z); T1.w = fmaf(w[i], a1.w, T1.w);
            T2.x = fmaf(w[i], a2.x, T2.x); T2.y = fmaf(w[i], a2.y, T2.y); T2.z = fmaf(w[i], a2.z, T2.z); T2.w = fmaf(w[i], a2.w, T2.w);
        }
        const float* vp = v_posed + (size_t)b * ld_vp + (size_t)v * 3;
        const float x = vp[0], y = vp[1], z = vp[2];
        float ox = ((T0.x * x + T0.y * y) + T0.z * z) + T0.w;
        float oy = ((T1.x * x + T1.y * y) + T1.z * z) + T1.w;
        float oz = ((T2.x * x + T2.y * y) + T2.z * z) + T2.w;
        if (trans) { ox += trans[(size_t)b * 3]; oy += trans[(size_t)b * 3 + 1]; oz += trans[(size_t)b * 3 + 2]; }
        float* o = verts + ((size_t)b * V + v) * 3;
        o[0] = ox * out_scale; o[1] = oy * out_scale; o[2] = oz * out_scale;
    }
}

// Sparse skinning: every SMPL vertex is bound to at most FOUR joints (the shipped th_weights have <= 4 non-zeros per row), so
// T_v = sum over the 4 (joint, weight) pairs of the vertex (ascending joint order: the same sum as the dense one, whose other 20
// terms are exact zeros). 12 shared-memory loads + 48 FMAs per vertex and sample instead of 72 + 288: the kernel becomes
// what it should be, a stream over v_posed / verts. Same CTA shape as the dense kernel (SMPL_SKIN_NB samples per CTA).
__global__ void __launch_bounds__(256)
smpl_skin4_kernel(const float* __restrict__ v_posed, const float* __restrict__ Amat, const int4* __restrict__ idx4, const float4* __restrict__ w4,
                  const float* __restrict__ trans, int V, int B, int ld_vp, float* __restrict__ verts, float out_scale) {
    // thread = (a PAIR of consecutive vertices, one sample): the pair's 6 coordinates are three aligned 8-byte words in v_posed
    // (row stride ld_vp, even) and in verts (V * 3 floats per sample with V even), so all global traffic is float2; the sparse
    // table (32 B per vertex) is re-read per sample from L2 - cheap next to what the dense kernel's sample loop saved.
    // The sample's 24 transforms sit in shared memory TRANSPOSED, As[k][joint] (k = 0..11): a warp's lanes read element k of
    // 32 different joints from 32 different banks (joint < 24 < 32), whatever the joints are - the natural [joint][12] layout read
    // with 16-byte loads had 42 % of its wavefronts in bank conflicts on vertices whose neighbours use unrelated joints
    __shared__ float As[12 * 24];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 288; i += blockDim.x) As[(i % 12) * 24 + i / 12] = Amat[(size_t)b * 288 + i];
    __syncthreads();
    const int v = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (v >= V) return;
    const bool two = v + 1 < V;
    const int4 ia = idx4[v], ib = two ? idx4[v + 1] : ia;
    const float4 wa = w4[v], wb = two ? w4[v + 1] : wa;
    const int ids[2][4] = {{ia.x, ia.y, ia.z, ia.w}, {ib.x, ib.y, ib.z, ib.w}};
    const float ws[2][4] = {{wa.x, wa.y, wa.z, wa.w}, {wb.x, wb.y, wb.z, wb.w}};
    const float2* vp = reinterpret_cast<const float2*>(v_posed + (size_t)b * ld_vp + (size_t)v * 3);
    float2 p0 = vp[0], p1 = make_float2(0.f, 0.f), p2 = p1;
    if (two) { p1 = vp[1]; p2 = vp[2]; } else { p1.x = v_posed[(size_t)b * ld_vp + (size_t)v * 3 + 2]; }
    const float xyz[2][3] = {{p0.x, p0.y, p1.x}, {p1.y, p2.x, p2.y}};
    float o[2][3];
    const float tx = trans ? trans[(size_t)b * 3] : 0.f, ty = trans ? trans[(size_t)b * 3 + 1] : 0.f, tz = trans ? trans[(size_t)b * 3 + 2] : 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float w = ws[q][i];
            const float* a = As + ids[q][i];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = fmaf(w, a[k * 24], T[k]);
        }
        const float x = xyz[q][0], y = xyz[q][1], z = xyz[q][2];
        float ox = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        float oy = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        float oz = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
        if (trans) { ox += tx; oy += ty; oz += tz; }
        o[q][0] = ox * out_scale; o[q][1] = oy * out_scale; o[q][2] = oz * out_scale;
    }
    float* op = verts + ((size_t)b * V + v) * 3;
    if (two && (V & 1) == 0) {
        float2* o2 = reinterpret_cast<float2*>(op);
        o2[0] = make_float2(o[0][0], o[0][1]); o2[1] = make_float2(o[0][2], o[1][0]); o2[2] = make_float2(o[1][1], o[1][2]);
    } else {
        op[0] = o[0][0]; op[1] = o[0][1]; op[2] = o[0][2];
        if (two) { op[3] = o[1][0]; op[4] = o[1][1]; op[5] = o[1][2]; }
    }
}
