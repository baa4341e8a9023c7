// libpmce_b200: C-ABI entry points and host-side orchestration of the PMCE forward hot path.
// All device work is enqueued on the caller's stream; no allocation, no synchronisation (graph-capturable).
//
// Data flow: the residual streams and everything attention reads stay fp32; every tensor that is the A operand of a
// projection is produced directly in split-bf16 form (hi/lo pair) by the kernel that computes it, and every weight
// matrix has a split-bf16 copy (made once by pmce_pack_weights), so all projections run on the tcgen05 GEMM
// (gemm_tc.cuh). The fp32 CUDA-core GEMM (gemm_simt.cuh) remains as the exact reference entry point (pmce_linear).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <nvtx3/nvToolsExt.h>
#include "../../include/pmce_b200.h"
#include "layout.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "gru_tc.cuh"
#include "attn_tc.cuh"
#include "attn_flash_tc.cuh"
#include "ca_fused.cuh"
#include "mlp_fused.cuh"
#include "attn_rows_tc.cuh"
#include "spin.cuh"

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            pmce_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 10;                                                                             \
        }                                                                                          \
    } while (0)
#define CKL() do { count_launch(); CK(cudaGetLastError()); } while (0)
#define CKG(expr) do { count_launch(); CK(expr); } while (0)
// kernel launch with programmatic stream serialisation (host_once.h pmce_launch): the kernel calls pdl_wait() / pdl_enter()
#define PLAUNCH(kernel, grid, block, smem, st, ...) CKG(pmce_launch(kernel, dim3(grid), dim3(block), smem, st, 0, __VA_ARGS__))
#define RET(x) do { int _r = (x); if (_r) return _r; } while (0)

// NVTX ranges (header-only NVTX3; a no-op unless a profiler is attached) around the stages of the forward, so nsys / ncu
// --nvtx timelines read "pmce/lifter", "pmce/image-feature stream", "pmce/decoder" ... instead of 100+ anonymous launches.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static std::atomic<unsigned long long> g_launches{0};
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long pmce_launch_count(void) { return g_launches.load(); }

namespace {

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------------
// workspace carving (256-byte aligned chunks)
// ---------------------------------------------------------------------------------------------------
struct Carver {
    char* base;
    size_t cur;   // bytes
    void* take_bytes(size_t n) {
        void* p = base ? base + cur : nullptr;
        cur += (n + 255) / 256 * 256;
        return p;
    }
    float* f32(size_t n) { return (float*)take_bytes(n * 4); }
    SplitOut split(size_t n) {
        SplitOut s;
        s.hi = (bf16*)take_bytes(n * 2);
        s.lo = (bf16*)take_bytes(n * 2);
        return s;
    }
};

struct Workspace {
    SplitOut feat_s;                                   // img_feat [B*T, F] split (shared by lifter embed and GRU)
    // lifter
    float *imgemb, *x, *qkv, *r3, *joints_m;
    SplitOut xn_s, att_s, hid_s;
    // gru
    float *gi0, *y0, *gi1f, *gi1b, *h1[2][2], *g, *y1;
    SplitOut y0_s, g_s, gr_s, h1_s[2][2], y1_s;
    unsigned* gru_counters;                            // step-barrier counters of the persistent GRU layer kernel (2 layers x 2 directions)
    // decoder
    float *gb, *verts[3], *Jf, *Vf, *xqv, *xkj, *Kj, *Vj, *qkv_d;
    float *xqj, *xkv, *Kv, *Vv, *qkvj;
    SplitOut Jf_s, Vf_s, tA_s, tA2_s, tB_s, tJ_s, tJ2_s, att_ds, hid_ds, attj_s, hidj_s, im2col_s;
    float* lc_mesh;
    CaFolded fold[3];                                  // per-clip folded operands of the fused vertex cross-attention, per block
    float* etab[3];                                    // per block: E = b_e + vpos + vQ [Vd, 64], the embed-mode table of that kernel
    size_t bytes;
};

Workspace carve(const pmce_dims_t& d, int B, void* base) {
    Workspace w;
    Carver c{(char*)base, 0};
    const size_t J = d.num_joint, C = d.embed_dim, T = d.seqlen, Vd = d.num_vert_ds, H = d.gru_hidden, F = d.feat_dim, D = d.coevo_dim;
    const size_t N = (size_t)B * T * J;
    w.feat_s = c.split((size_t)B * T * F);
    w.imgemb = c.f32((size_t)B * T * C);
    w.x = c.f32(N * C); w.qkv = c.f32(N * 3 * C); w.r3 = c.f32(N * 3); w.joints_m = c.f32((size_t)B * J * 3);
    w.xn_s = c.split(N * C); w.att_s = c.split(N * C); w.hid_s = c.split(N * 2 * C);
    const size_t mid = T / 2;
    w.gi0 = c.f32((size_t)T * B * 6 * H); w.y0 = c.f32((size_t)T * B * 2 * H);
    w.gi1f = c.f32((mid + 1) * B * 3 * H); w.gi1b = c.f32((T - mid) * B * 3 * H);
    for (int dir = 0; dir < 2; ++dir) for (int i = 0; i < 2; ++i) w.h1[dir][i] = c.f32((size_t)B * H);
    w.g = c.f32((size_t)B * F);
    w.y0_s = c.split((size_t)T * B * 2 * H); w.g_s = c.split((size_t)B * F); w.gr_s = c.split((size_t)B * F);
    for (int dir = 0; dir < 2; ++dir) for (int i = 0; i < 2; ++i) w.h1_s[dir][i] = c.split((size_t)B * H);
    w.y1 = c.f32((size_t)(T + 1) * B * H); w.y1_s = c.split((size_t)(T + 1) * B * H);     // layer-1 hidden states of the (T/2+1) + (T-T/2) live steps
    w.gru_counters = (unsigned*)c.take_bytes(2 * 2 * GRU_MAX_CTAS * sizeof(unsigned));     // step flags: 2 layers x 2 directions x 64 CTAs
    w.gb = c.f32((size_t)B * PMCE_ADALN_SLOTS * 2 * D);
    for (int i = 0; i < 3; ++i) w.verts[i] = c.f32((size_t)B * Vd * 3);
    const size_t nv = (size_t)B * Vd, nj = (size_t)B * J;
    w.Jf = c.f32(nj * D); w.Vf = c.f32(nv * D); w.xqv = c.f32(nv * D);
    w.xkj = c.f32(nj * D); w.Kj = c.f32(nj * D); w.Vj = c.f32(nj * D); w.qkv_d = c.f32(nv * 4 * D);      /* fp32 qkv [nv,192] or the bf16 [nv,512] attention records */
    w.xqj = c.f32(nj * D); w.xkv = c.f32(nv * D); w.Kv = c.f32(nv * D); w.Vv = c.f32(nv * D); w.qkvj = c.f32(nj * 3 * D);
    w.Jf_s = c.split(nj * D); w.Vf_s = c.split(nv * D); w.tA_s = c.split(nv * D); w.tA2_s = c.split(nv * D); w.tJ_s = c.split(nj * D);
    w.att_ds = c.split(nv * D); w.hid_ds = c.split(nv * 4 * D); w.attj_s = c.split(nj * D); w.hidj_s = c.split(nj * 4 * D);
    w.im2col_s = c.split((size_t)B * 3 * ((Vd * 3 + 7) / 8 * 8));
    w.tB_s = c.split(nv * D); w.tJ2_s = c.split(nj * D);
    w.lc_mesh = c.f32((size_t)B * d.num_vert * 3);
    for (int k = 0; k < 3; ++k) {
        SplitOut kq = c.split((size_t)B * CAF_NS * 64), vp = c.split((size_t)B * CAF_NS * 64);
        w.fold[k].kq_hi = kq.hi; w.fold[k].kq_lo = kq.lo; w.fold[k].vp_hi = vp.hi; w.fold[k].vp_lo = vp.lo;
        w.fold[k].sb = c.f32((size_t)B * CAF_NS);
    }
    for (int k = 0; k < 3; ++k) w.etab[k] = c.f32(Vd * D);
    w.bytes = c.cur;
    return w;
}

int check_ws(const pmce_dims_t& d, int B, void* ws, size_t bytes, Workspace* out) {
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    if (!ws) { pmce_set_error("workspace is NULL"); return 2; }
    if (((uintptr_t)ws) & 255) { pmce_set_error("workspace must be 256-byte aligned"); return 2; }
    *out = carve(d, B, ws);
    if (out->bytes > bytes) { pmce_set_error("workspace too small: need %zu bytes, got %zu", out->bytes, bytes); return 2; }
    return 0;
}

// packed weights: fp32 region, then bf16 hi copy, then bf16 lo copy (same element offsets)
struct Weights {
    const float* f;
    const bf16* hi;
    const bf16* lo;
};
Weights weights_view(const Layout& L, const void* blob) {
    Weights w;
    w.f = (const float*)blob;
    w.hi = (const bf16*)(w.f + L.total_floats);
    w.lo = w.hi + L.total_floats;
    return w;
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
struct EpiOpt {
    const float* bias = nullptr;
    int act = 0;
    const float* resid = nullptr; int ld_resid = 0;
    const float* rowadd = nullptr; int period = 1;
    float* out = nullptr; int ld_out = 0;
    SplitOut outs{nullptr, nullptr}; int ld_split = 0;
    bf16* att = nullptr; float qscale = 1.0f;       // TC_ATTN32: operand records of attn_rows_tc_kernel (gemm_tc.cuh)
    bool mapped = false; RowMap rmap{1, 0, 0}, cmap{1, 0, 0};
    int force_bn = 0;                                // tile width of this launch (0: cost model)
};

// out = epi(A[M,K] * W[N,K]^T) on the tcgen05 path; A split [M,K] (ld lda), W at float offset w_off in the blob (ld ldw).
int linear_tc(const SplitOut& A, int lda, int M, int K, const Weights& W, size_t w_off, int ldw, int N, const EpiOpt& o, cudaStream_t st) {
    if ((K & 7) || (lda & 7) || (ldw & 7)) { pmce_set_error("linear_tc: K/lda/ldw must be multiples of 8 (K=%d lda=%d ldw=%d)", K, lda, ldw); return 3; }
    TcOperand a{A.hi, A.lo, M, K, lda}, w{W.hi + w_off, W.lo + w_off, N, K, ldw};
    TcEpi e;
    memset(&e, 0, sizeof(e));
    e.bias = o.bias; e.act = o.act; e.resid = o.resid; e.ld_resid = o.ld_resid; e.rowadd = o.rowadd; e.rowadd_period = o.period;
    e.out_f32 = o.out; e.ld_out = o.ld_out; e.out_hi = o.outs.hi; e.out_lo = o.outs.lo; e.ld_split = o.ld_split;
    e.mapped = o.mapped ? 1 : 0; e.rmap = o.rmap; e.cmap = o.cmap;
    e.out_att = o.att; e.qscale = o.qscale; e.force_bn = o.force_bn;
    count_launch();
    const int rc = launch_linear_tc(a, w, e, st);
    if (rc) { pmce_set_error("tensor-core GEMM launch failed (%d; M=%d N=%d K=%d): %s", rc, M, N, K, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

int split_rows(const float* x, int rows, int cols, int ld, bool relu, const SplitOut& o, int ld_out, cudaStream_t st) {
    PLAUNCH(split_rows_kernel, cdiv((long long)rows * (cols / 4), 256), 256, 0, st, x, rows, cols, ld, relu ? 1 : 0, o.hi, o.lo, ld_out);
    return 0;
}

template <int D>
int launch_attn_d(const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, float* O, SplitOut Os, AttnAddr ao, int nseq,
                  int H, int N1, int N2, cudaStream_t st) {
    const int fewq = pmce_env_int("PMCE_ATTN_FEWQ", 1);     // 0: the lane-per-query kernel for every shape (A/B, tests; read live)
    if (fewq && N1 <= 32 && N2 >= 64 && N2 <= 32 * ATTN_FEWQ_KPL && nseq <= 65535) {
        // few queries over many keys (joint <- vertex cross-attention): keys across the lanes, a warp per query
        const size_t sm = (size_t)N2 * (D + 4) * 2 * sizeof(float);
        if (sm <= 227 * 1024) {
            if (sm > 48 * 1024 && !pmce_configure_smem<attn_fewq_kernel<D>>(227 * 1024)) { pmce_set_error("attn_fewq: cudaFuncSetAttribute failed"); return 10; }
            PLAUNCH(attn_fewq_kernel<D>, dim3(H, nseq), 256, sm, st, Q, aq, K, V, akv, O, Os, ao, N1, N2, 1.0f / sqrtf((float)D));
            return 0;
        }
    }
    const size_t smem = (size_t)N2 * D * 2 * sizeof(float);
    if (smem > 227 * 1024) { pmce_set_error("attention K/V tile (%zu B) exceeds shared memory", smem); return 3; }
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int threads = N1 >= 128 ? 128 : (N1 > 64 ? 96 : (N1 > 32 ? 64 : 32));
    if (nseq > 65535) { pmce_set_error("too many attention sequences (%d) for one launch", nseq); return 3; }
    dim3 grid(cdiv(N1, threads), H, nseq);
    PLAUNCH(attn_kernel<D>, grid, threads, smem, st, Q, aq, K, V, akv, O, Os, ao, N1, N2, 1.0f / sqrtf((float)D));
    return 0;
}

int launch_attn(int D, const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, float* O, SplitOut Os, AttnAddr ao,
                int nseq, int H, int N1, int N2, cudaStream_t st) {
    switch (D) {
        case 8: return launch_attn_d<8>(Q, aq, K, V, akv, O, Os, ao, nseq, H, N1, N2, st);
        case 16: return launch_attn_d<16>(Q, aq, K, V, akv, O, Os, ao, nseq, H, N1, N2, st);
        case 32: return launch_attn_d<32>(Q, aq, K, V, akv, O, Os, ao, nseq, H, N1, N2, st);
        case 64: return launch_attn_d<64>(Q, aq, K, V, akv, O, Os, ao, nseq, H, N1, N2, st);
    }
    pmce_set_error("unsupported head_dim %d", D);
    return 3;
}

AttnAddr addr_plain(int ntok, int ld) {
    AttnAddr a;
    a.seq.div = 1; a.seq.s0 = ntok; a.seq.s1 = 0; a.tok = 1; a.ld = ld;
    return a;
}

const SplitOut NO_SPLIT{nullptr, nullptr};

// head_dim-32 attention on tcgen05 (attn_flash_tc.cuh): the decoder's vertex self- and cross-attention
int flash_attn32(const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, const SplitOut& Os, AttnAddr ao, int nseq, int H, int N1,
                 int N2, cudaStream_t st) {
    count_launch();
    const int rc = launch_attn_flash_tc(Q, aq, K, V, akv, Os, ao, nseq, H, N1, N2, st);
    if (rc) { pmce_set_error("attn_flash_tc launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

int ln_rows(const float* x, int nrows, int C, const LnParams* a, const float* pos, int pos_div, int pos_mod, float* out1,
            const LnParams* b, const SplitOut& out2s, cudaStream_t st, int map_rows = 0, int map_stride = 0) {
    LnParams za{nullptr, nullptr, 0.f};
    const int nv = C / 128;
    const dim3 grid(cdiv(nrows, 8));
#define LN_LAUNCH(MV) PLAUNCH(ln_rows_kernel<MV>, grid, 256, 0, st, x, nrows, C, a ? *a : za, a ? 1 : 0, pos, pos_div, pos_mod, out1, b ? *b : za, nullptr, \
                              b ? out2s : NO_SPLIT, map_rows, map_stride)
    if (nv <= 1) LN_LAUNCH(1); else if (nv <= 2) LN_LAUNCH(2); else if (nv <= 4) LN_LAUNCH(4); else LN_LAUNCH(8);
#undef LN_LAUNCH
    return 0;
}

int adaln(const float* x, int B, int ntok, const float* gb, int slot, const SplitOut& ys, cudaStream_t st) {
    const int nrows = B * ntok;
    PLAUNCH(adaln_apply_kernel, cdiv(nrows, 8), 256, 0, st, x, nrows, ntok, gb, PMCE_ADALN_SLOTS * 128, slot, 1e-6f, nullptr, ys);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a2/a3 lifter
// ---------------------------------------------------------------------------------------------------
// F = number of frames the token buffer holds (F * J tokens): B * T for window-major batches; the per-frame spatial pass of
// pmce_forward_sliding runs on the distinct frames of overlapping windows. The temporal pass needs window-major tokens (F % T == 0).
int vit_block(const pmce_dims_t& d, const Weights& W, const VitBlockW& w, const Workspace& ws, int F, bool temporal, cudaStream_t st) {
    const int J = d.num_joint, C = d.embed_dim, T = d.seqlen, Hh = d.lifter_heads;
    const int N = F * J;
    const int B = F / T;
    const bool tc_attn = (C / Hh == 64) && (temporal ? (T <= 128 && T % 8 == 0) : (J <= 128));
    // split-bf16 view of the qkv buffer (same bytes as the fp32 one): written by the projection when the tensor-core
    // attention consumes it by TMA
    SplitOut qkv_s{reinterpret_cast<bf16*>(ws.qkv), reinterpret_cast<bf16*>(ws.qkv) + (size_t)N * 3 * C};
    {   // qkv = W_qkv LN1(x) + b
        EpiOpt o; o.bias = W.f + w.qkvb;
        if (tc_attn) { o.outs = qkv_s; o.ld_split = 3 * C; } else { o.out = ws.qkv; o.ld_out = 3 * C; }
        RET(linear_tc(ws.xn_s, C, N, C, W, w.qkvw, C, 3 * C, o, st));
    }
    AttnAddr a, ao;
    int nseq, L;
    if (!temporal) { a.seq.div = 1; a.seq.s0 = J; a.seq.s1 = 0; a.tok = 1; nseq = F; L = J; }
    else { a.seq.div = J; a.seq.s0 = (long long)T * J; a.seq.s1 = 1; a.tok = J; nseq = B * J; L = T; }
    a.ld = 3 * C;
    ao = a; ao.ld = C;
    if (tc_attn) {                       // head_dim 64: packed block-diagonal attention on tcgen05 (attn_tc.cuh)
        count_launch();
        const int rc = launch_attn_tile_tc(qkv_s.hi, qkv_s.lo, N, C, J, T, temporal, ws.att_s, nseq, Hh, st);
        if (rc) { pmce_set_error("attn_tile_tc launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    } else {
    const int chunk = (65535 / J) * J;   // gridDim.z limit; multiples of J keep the (b,j) decomposition intact
    for (int s0 = 0; s0 < nseq; s0 += chunk) {
        const int ns = nseq - s0 < chunk ? nseq - s0 : chunk;
        const long long r0 = a.seq(s0);
        SplitOut os{ws.att_s.hi + r0 * C, ws.att_s.lo + r0 * C};
        RET(launch_attn(C / Hh, ws.qkv + r0 * 3 * C, a, ws.qkv + r0 * 3 * C + C, ws.qkv + r0 * 3 * C + 2 * C, a, nullptr, os, ao, ns, Hh, L, L, st));
    }
    }
    {   // x += W_proj att + b
        EpiOpt o; o.bias = W.f + w.projb; o.resid = ws.x; o.ld_resid = C; o.out = ws.x; o.ld_out = C;
        RET(linear_tc(ws.att_s, C, N, C, W, w.projw, C, C, o, st));
    }
    LnParams n2{W.f + w.n2w, W.f + w.n2b, 1e-6f};
    RET(ln_rows(ws.x, N, C, nullptr, nullptr, 1, 1, nullptr, &n2, ws.xn_s, st));
    {   // hid = gelu(W_fc1 LN2(x) + b)
        EpiOpt o; o.bias = W.f + w.fc1b; o.act = 1; o.outs = ws.hid_s; o.ld_split = 2 * C;
        RET(linear_tc(ws.xn_s, C, N, C, W, w.fc1w, C, 2 * C, o, st));
    }
    {   // x += W_fc2 hid + b
        EpiOpt o; o.bias = W.f + w.fc2b; o.resid = ws.x; o.ld_resid = C; o.out = ws.x; o.ld_out = C;
        RET(linear_tc(ws.hid_s, 2 * C, N, 2 * C, W, w.fc2w, 2 * C, C, o, st));
    }
    return 0;
}

// Windows b = 0..B-1 of T frames each over a track of nfr frames: window b covers frames [b*fstride, b*fstride + T).
// fstride == T, nfr == B*T: the ordinary batch of materialised clips. fstride < T (pmce_forward_sliding): overlapping windows
// of one track - everything that is a function of a single frame (imgfeat_embed, the token embedding and SpatialBlocks[0],
// PoseEstimation.py:78-84: spatial attention only mixes the joints of one frame) is computed once per FRAME, then gathered
// into the window-major token order by the LayerNorm that follows it.
int lifter(const Layout& L, const Weights& W, const float* pose2d, int B, int nfr, int fstride, float* pose3d, const Workspace& ws,
           cudaStream_t st) {
    NvtxRange nvtx("pmce/lifter (pose stream)");
    const pmce_dims_t& d = L.d;
    const int J = d.num_joint, C = d.embed_dim, T = d.seqlen, F = d.feat_dim;
    const int N = B * T * J;
    const bool sliding = fstride != T;
    {   // imgfeat_embed for every frame (feat_s prepared by the caller)
        EpiOpt o; o.bias = W.f + L.ieb; o.out = ws.imgemb; o.ld_out = C;
        RET(linear_tc(ws.feat_s, F, nfr, F, W, L.iew, F, C, o, st));
    }
    LnParams s0n1{W.f + L.sp[0].n1w, W.f + L.sp[0].n1b, 1e-6f};
    const int nvc = C / 128;
#define ROWK_LAUNCH(KERNEL, GRID, ...)                                            \
    do {                                                                           \
        if (nvc <= 1) PLAUNCH(KERNEL<1>, GRID, 256, 0, st, __VA_ARGS__);           \
        else if (nvc <= 2) PLAUNCH(KERNEL<2>, GRID, 256, 0, st, __VA_ARGS__);      \
        else if (nvc <= 4) PLAUNCH(KERNEL<4>, GRID, 256, 0, st, __VA_ARGS__);      \
        else PLAUNCH(KERNEL<8>, GRID, 256, 0, st, __VA_ARGS__);                    \
    } while (0)
    ROWK_LAUNCH(lifter_embed_kernel, cdiv(nfr * J, 8), pose2d, ws.imgemb, W.f + L.jew, W.f + L.jeb, W.f + L.spos, nfr * J, J, C, s0n1, ws.x, nullptr, ws.xn_s);
    LnParams ns{W.f + L.nsw, W.f + L.nsb, 1e-6f}, nt{W.f + L.ntw, W.f + L.ntb, 1e-6f};
    for (int i = 0; i < d.depth; ++i) {
        RET(vit_block(d, W, L.sp[i], ws, i == 0 ? nfr : B * T, false, st));
        LnParams tn1{W.f + L.tp[i].n1w, W.f + L.tp[i].n1b, 1e-6f};
        if (i == 0 && sliding) {
            // frame-major tokens [nfr*J, C] -> window-major [B*T*J, C] while normalising; the copy keeps the source apart from
            // the destination (the qkv buffer is free until the next projection)
            CK(cudaMemcpyAsync(ws.qkv, ws.x, (size_t)nfr * J * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
            RET(ln_rows(ws.qkv, N, C, &ns, W.f + L.tpos, J, T, ws.x, &tn1, ws.xn_s, st, T * J, fstride * J));
        } else {
            RET(ln_rows(ws.x, N, C, &ns, i == 0 ? W.f + L.tpos : nullptr, J, T, ws.x, &tn1, ws.xn_s, st));
        }
        RET(vit_block(d, W, L.tp[i], ws, B * T, true, st));
        if (i + 1 < d.depth) {
            LnParams sn1{W.f + L.sp[i + 1].n1w, W.f + L.sp[i + 1].n1b, 1e-6f};
            RET(ln_rows(ws.x, N, C, &nt, nullptr, 1, 1, ws.x, &sn1, ws.xn_s, st));
        }
    }
    LnParams nh{W.f + L.r0w, W.f + L.r0b, 1e-5f};
    ROWK_LAUNCH(lifter_head_kernel, cdiv(N, 8), ws.x, N, C, nt, nh, W.f + L.r1w, W.f + L.r1b, ws.r3);
#undef ROWK_LAUNCH
    PLAUNCH(lifter_fuse_kernel, cdiv(B * J * 3, 256), 256, 0, st, ws.r3, W.f + L.fusw, W.f + L.fusb, B, T, J, pose3d, ws.joints_m);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a4 GRU
// ---------------------------------------------------------------------------------------------------
// One recurrent step for up to two directions. The first step of a direction (h_prev = 0) needs no GEMM and runs the
// CUDA-core kernel; every other step runs the tcgen05 kernel (gru_tc.cuh) on the split-bf16 copy of h_prev.
struct GruStep {
    GruDir d;            // fp32 view (gi, hprev, bhh, hout, hs, strides); d.whh = fp32 W_hh (CUDA-core path)
    SplitOut hprev_s;    // split copy of h_prev [B,H] (ld_hs) — null on the first step
    int ld_hs;
    size_t whh_off;      // float offset of W_hh [3H,H] in the weight blob
};

// Few-CTA recurrent steps (gru_step_multi_kernel, gru_tc.cuh) while the image-feature stream runs beside the pose lifter: an
// EXPERIMENT that did not pay, kept behind knobs (default off). Measured at 64 clips (tools/forward_breakdown.py, overlap_probe.py):
// the forward is 2,547 us, 2,116 us without the image-feature stream (580 us alone): only ~150 us of it hide under the 1,585 us
// lifter. Hypothesis: a 128-CTA step delays the lifter's persistent GEMM (static tile schedule) it collides with by the whole
// step, so run the first steps on 12 CTAs = the 6 TPCs the balanced GEMM grids leave free (tc_balanced_grid / tc_sm_reserve).
// Result: a 12-CTA step takes ~60-70 us (per-SM L2 ingest ~64 GB/s x 3.8 MB of W_hh + h per CTA) and costs the lifter as much
// as the 128-CTA step it replaces (forward 2,538 / 2,549 / 2,627 / 2,780 us with 0 / 12 / 19 / 22 such steps): the GEMMs are bound
// by L2 operand bandwidth and a step moves ~50 MB through L2 however many SMs it uses.
//   PMCE_GRU_FEW_STEPS  how many of the first recurrent steps run on few CTAs (default 0; -1 = a host-side estimate of what fits
//                       under the lifter)
//   PMCE_GRU_FEW        CTAs of such a step (default 12)
struct GruFew { int ctas, steps; };
GruFew gru_few_plan(const pmce_dims_t& d, int B) {
    static int ctas = -1, steps = -2;
    if (ctas < 0) { ctas = pmce_env_int("PMCE_GRU_FEW", 12) & ~1; steps = pmce_env_int("PMCE_GRU_FEW_STEPS", 0); }
    GruFew f{ctas, steps};
    if (ctas <= 0) { f.steps = 0; return f; }
    if (steps < 0) {
        const double c = d.embed_dim / 512.0, h = d.gru_hidden / 1024.0;
        const double lifter_us = 1590.0 * ((double)B * d.seqlen * d.num_joint / 17408.0) * c * c * (d.depth / 3.0);
        const double step_us = 60.0 * h * h * (double)cdiv(B, 128) * (12.0 / ctas);
        const double room = 0.85 * lifter_us - 200.0;      // the stream's GEMMs (input projections) take ~200 us
        f.steps = room > 0 ? (int)(room / step_us) : 0;
        if (f.steps < 4) f.steps = 0;
    }
    return f;
}

template <int U, int S>
int launch_gru_multi(const GruStep* s, int ndir, const Weights& W, int B, int H, int ctas, cudaStream_t st) {
    using Cfg = GruMultiCfg<U, S>;
    GruTcMaps maps[2];
    GruTcDir dirs[2];
    for (int i = 0; i < 2; ++i) {
        const GruStep& x = s[i < ndir ? i : 0];
        if (make_tmap_bf16(&maps[i].h_hi, x.hprev_s.hi, B, H, x.ld_hs, 128) || make_tmap_bf16(&maps[i].h_lo, x.hprev_s.lo, B, H, x.ld_hs, 128) ||
            make_tmap_bf16(&maps[i].w_hi, W.hi + x.whh_off, 3 * H, H, H, U) || make_tmap_bf16(&maps[i].w_lo, W.lo + x.whh_off, 3 * H, H, H, U)) {
            pmce_set_error("gru_step: cuTensorMapEncodeTiled failed");
            return 10;
        }
        dirs[i].gi = x.d.gi; dirs[i].hprev = x.d.hprev; dirs[i].bhh = x.d.bhh; dirs[i].hout = x.d.hout; dirs[i].hs = x.d.hs;
        dirs[i].ld_gi = x.d.ld_gi; dirs[i].ld_h = x.d.ld_h; dirs[i].ld_o = x.d.ld_o; dirs[i].ld_s = x.d.ld_s;
    }
    if (!pmce_configure_smem<gru_step_multi_kernel<U, S>>(Cfg::SMEM)) { pmce_set_error("gru_step: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); return 10; }
    const int ntiles = (H / U) * ndir * cdiv(B, 128);
    int grid = ctas < ntiles ? ctas : ntiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (grid >= 2) {       // CTA pairs land on whole TPCs, so the lifter's pair GEMMs find their TPCs whole as well
        grid &= ~1;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        nattr = 1;
    }
    cfg.gridDim = dim3(grid); cfg.attrs = attr; cfg.numAttrs = nattr;
    count_launch();
    if (cudaLaunchKernelEx(&cfg, gru_step_multi_kernel<U, S>, maps[0], maps[1], dirs[0], dirs[1], B, H, ndir) != cudaSuccess) {
        pmce_set_error("gru_step_multi launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 10;
    }
    return 0;
}

// PMCE_SKIP (profiling build only, results are garbage): bit 0 = no recurrent GRU steps, bit 1 = no image-feature stream at all,
// bit 2 = no lifter, bit 3 = no decoder - tools/forward_breakdown.py times the forward with parts removed
static int skip_mask() {
    static int m = -1;
    if (m < 0) m = pmce_profiling_knob("PMCE_SKIP");
    return m;
}

int gru_step(const GruStep* s, int ndir, const Weights& W, int B, int H, int few, cudaStream_t st) {
    if (skip_mask() & 1) return 0;
    if (s[0].d.hprev && few > 0 && H % 64 == 0) {
        static int u = -1;
        if (u < 0) u = pmce_env_int("PMCE_GRU_FEW_U", 32);
        return u == 64 ? launch_gru_multi<64, 2>(s, ndir, W, B, H, few, st) : launch_gru_multi<32, 4>(s, ndir, W, B, H, few, st);
    }
    if (!s[0].d.hprev) {   // first step of both directions (they always start together)
        dim3 grid(H / 16, cdiv(B, 64), ndir);
        PLAUNCH(gru_step_kernel, grid, 256, 0, st, s[0].d, s[ndir > 1 ? 1 : 0].d, B, H);
        return 0;
    }
    GruTcMaps maps[2];
    GruTcDir dirs[2];
    for (int i = 0; i < 2; ++i) {
        const GruStep& x = s[i < ndir ? i : 0];
        if (make_tmap_bf16(&maps[i].h_hi, x.hprev_s.hi, B, H, x.ld_hs, 128) || make_tmap_bf16(&maps[i].h_lo, x.hprev_s.lo, B, H, x.ld_hs, 128) ||
            make_tmap_bf16(&maps[i].w_hi, W.hi + x.whh_off, 3 * H, H, H, GRU_U) || make_tmap_bf16(&maps[i].w_lo, W.lo + x.whh_off, 3 * H, H, H, GRU_U)) {
            pmce_set_error("gru_step: cuTensorMapEncodeTiled failed");
            return 10;
        }
        dirs[i].gi = x.d.gi; dirs[i].hprev = x.d.hprev; dirs[i].bhh = x.d.bhh; dirs[i].hout = x.d.hout; dirs[i].hs = x.d.hs;
        dirs[i].ld_gi = x.d.ld_gi; dirs[i].ld_h = x.d.ld_h; dirs[i].ld_o = x.d.ld_o; dirs[i].ld_s = x.d.ld_s;
    }
    if (!pmce_configure_smem<gru_step_tc_kernel>(GRU_SMEM)) { pmce_set_error("gru_step: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); return 10; }
    dim3 grid(H / GRU_U, ndir, cdiv(B, 128));
    PLAUNCH(gru_step_tc_kernel, grid, 192, GRU_SMEM, st, maps[0], maps[1], dirs[0], dirs[1], B, H);
    return 0;
}

bool gru_persistent_enabled() {
    // PMCE_GRU_PERSISTENT=1: one persistent launch per GRU layer (gru_layer_tc_kernel) instead of one launch per time step.
    // Measured (B=64, same box): gru_mid 583 us persistent vs 514 us per step, whole forward 24.3k vs 25.3k clips/s - a step is
    // bound by what ONE SM must ingest from L2 (its 192 KB W_hh slice + the whole 256 KB h_prev, ~8 us at ~57 GB/s per SM), not by
    // the launch; the persistent form adds the step barrier's release/poll latency on top and pins 128 SMs. Off by default.
    static int on = -1;
    if (on < 0) on = pmce_env_int("PMCE_GRU_PERSISTENT", 0) ? 1 : 0;
    return on == 1;
}

// one persistent launch per batch tile of 128 rows for a whole GRU layer (gru_tc.cuh)
int gru_layer(const GruLayerDir* dd, int ndir, const SplitOut* hs, const int* nblk, const Weights& W, const size_t* whh_off, int B, int H,
              unsigned* counters, cudaStream_t st) {
    GruTcMaps maps[2];
    for (int i = 0; i < 2; ++i) {
        const int k = i < ndir ? i : 0;
        if (make_tmap_bf16_3d(&maps[i].h_hi, hs[k].hi, dd[k].ld_y, B, nblk[k], dd[k].ld_y, (long long)B * dd[k].ld_y, 128, 1) ||
            make_tmap_bf16_3d(&maps[i].h_lo, hs[k].lo, dd[k].ld_y, B, nblk[k], dd[k].ld_y, (long long)B * dd[k].ld_y, 128, 1) ||
            make_tmap_bf16(&maps[i].w_hi, W.hi + whh_off[k], 3 * H, H, H, GRU_U) || make_tmap_bf16(&maps[i].w_lo, W.lo + whh_off[k], 3 * H, H, H, GRU_U)) {
            pmce_set_error("gru_layer: cuTensorMapEncodeTiled failed");
            return 10;
        }
    }
    if (!pmce_configure_smem<gru_layer_tc_kernel>(GRU_SMEM)) { pmce_set_error("gru_layer: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); return 10; }
    for (int b0 = 0; b0 < B; b0 += 128) {          // all CTAs of a launch must be co-resident: one 128-row batch tile at a time
        CK(cudaMemsetAsync(counters, 0, 2 * GRU_MAX_CTAS * sizeof(unsigned), st));
        gru_layer_tc_kernel<<<dim3(H / GRU_U, ndir), 192, GRU_SMEM, st>>>(maps[0], maps[1], dd[0], dd[ndir > 1 ? 1 : 0], B, H, b0, counters);
        CKL();
    }
    return 0;
}

// nfr / fstride as in lifter(): the layer-0 input projection is a per-frame product, so overlapping windows share it.
int gru_mid(const Layout& L, const Weights& W, int B, int nfr, int fstride, float* g, const Workspace& ws, cudaStream_t st, GruFew few = GruFew{0, 0}) {
    const pmce_dims_t& d = L.d;
    const int T = d.seqlen, H = d.gru_hidden, F = d.feat_dim;
    const int mid = T / 2;
    {   // layer-0 input projections, every frame, both directions: gi0[frame][6H] (the step kernels stride over windows with fstride*6H)
        EpiOpt o; o.bias = W.f + L.bih0; o.out = ws.gi0; o.ld_out = 6 * H;
        RET(linear_tc(ws.feat_s, F, nfr, F, W, L.wih0, F, 6 * H, o, st));
    }
    if (gru_persistent_enabled() && H % 64 == 0 && H / GRU_U <= 64) {
        const int nf = mid + 1, nb = T - mid;
        {   // layer 0: y0 [T, B, 2H]; forward direction writes blocks 0..T-1, backward T-1..0 (block = frame)
            GruLayerDir dd[2];
            memset(dd, 0, sizeof(dd));
            for (int dir = 0; dir < 2; ++dir) {
                GruLayerDir& x = dd[dir];
                x.gi = ws.gi0 + (dir == 0 ? 0 : (size_t)(T - 1) * 6 * H) + dir * 3 * H;
                x.gi_step = dir == 0 ? 6 * H : -6 * H; x.ld_gi = fstride * 6 * H;
                x.y = ws.y0; x.ys = ws.y0_s; x.ld_y = 2 * H; x.col0 = dir * H;
                x.blk0 = dir == 0 ? 0 : T - 1; x.blk_step = dir == 0 ? 1 : -1; x.nsteps = T;
                x.bhh = W.f + L.bhh0[dir];
            }
            const SplitOut hs[2] = {ws.y0_s, ws.y0_s};
            const int nblk[2] = {T, T};
            const size_t wo[2] = {L.whh0[0], L.whh0[1]};
            RET(gru_layer(dd, 2, hs, nblk, W, wo, B, H, ws.gru_counters, st));
        }
        {
            EpiOpt o; o.bias = W.f + L.bih1[0]; o.out = ws.gi1f; o.ld_out = 3 * H;
            RET(linear_tc(ws.y0_s, 2 * H, nf * B, 2 * H, W, L.wih1[0], 2 * H, 3 * H, o, st));
        }
        {
            EpiOpt o; o.bias = W.f + L.bih1[1]; o.out = ws.gi1b; o.ld_out = 3 * H;
            SplitOut a{ws.y0_s.hi + (size_t)mid * B * 2 * H, ws.y0_s.lo + (size_t)mid * B * 2 * H};
            RET(linear_tc(a, 2 * H, nb * B, 2 * H, W, L.wih1[1], 2 * H, 3 * H, o, st));
        }
        {   // layer 1: only the steps y[T//2] depends on; y1 [nf + nb, B, H]: forward blocks 0..nf-1, backward nf..nf+nb-1
            GruLayerDir dd[2];
            memset(dd, 0, sizeof(dd));
            for (int dir = 0; dir < 2; ++dir) {
                GruLayerDir& x = dd[dir];
                x.gi = dir == 0 ? ws.gi1f : ws.gi1b + (size_t)(nb - 1) * B * 3 * H;      // gi1b row block i = frame mid + i; step s reads frame T-1-s
                x.gi_step = dir == 0 ? (long long)B * 3 * H : -(long long)B * 3 * H; x.ld_gi = 3 * H;
                x.y = ws.y1; x.ys = ws.y1_s; x.ld_y = H; x.col0 = 0;
                x.blk0 = dir == 0 ? 0 : nf; x.blk_step = 1; x.nsteps = dir == 0 ? nf : nb;
                x.bhh = W.f + L.bhh1[dir];
                x.last_out = g + dir * H; x.ld_last = 2 * H;
            }
            const SplitOut hs[2] = {ws.y1_s, ws.y1_s};
            const int nblk[2] = {nf + nb, nf + nb};
            const size_t wo[2] = {L.whh1[0], L.whh1[1]};
            RET(gru_layer(dd, 2, hs, nblk, W, wo, B, H, ws.gru_counters + 2 * GRU_MAX_CTAS, st));
        }
        return 0;
    }
    for (int s = 0; s < T; ++s) {
        GruStep dd[2];
        for (int dir = 0; dir < 2; ++dir) {
            const int t = dir == 0 ? s : T - 1 - s;            // frame this step produces
            const int tp = dir == 0 ? t - 1 : t + 1;           // frame h_prev belongs to
            const size_t off = (size_t)t * B * 2 * H + dir * H, offp = (size_t)tp * B * 2 * H + dir * H;
            GruStep& x = dd[dir];
            x.d.gi = ws.gi0 + (size_t)t * 6 * H + dir * 3 * H; x.d.ld_gi = fstride * 6 * H;
            x.d.hprev = s > 0 ? ws.y0 + offp : nullptr; x.d.ld_h = 2 * H;
            x.d.whh = W.f + L.whh0[dir]; x.d.bhh = W.f + L.bhh0[dir];
            x.d.hout = ws.y0 + off; x.d.ld_o = 2 * H;
            x.d.hs.hi = ws.y0_s.hi + off; x.d.hs.lo = ws.y0_s.lo + off; x.d.ld_s = 2 * H;
            x.hprev_s.hi = s > 0 ? ws.y0_s.hi + offp : nullptr; x.hprev_s.lo = s > 0 ? ws.y0_s.lo + offp : nullptr; x.ld_hs = 2 * H;
            x.whh_off = L.whh0[dir];
        }
        RET(gru_step(dd, 2, W, B, H, (s > 0 && few.steps-- > 0) ? few.ctas : 0, st));
    }
    // layer 1: only the steps y[T//2] depends on (fwd t = 0..mid, bwd t = T-1..mid)
    const int nf = mid + 1, nb = T - mid;
    {
        EpiOpt o; o.bias = W.f + L.bih1[0]; o.out = ws.gi1f; o.ld_out = 3 * H;
        RET(linear_tc(ws.y0_s, 2 * H, nf * B, 2 * H, W, L.wih1[0], 2 * H, 3 * H, o, st));
    }
    {
        EpiOpt o; o.bias = W.f + L.bih1[1]; o.out = ws.gi1b; o.ld_out = 3 * H;
        SplitOut a{ws.y0_s.hi + (size_t)mid * B * 2 * H, ws.y0_s.lo + (size_t)mid * B * 2 * H};
        RET(linear_tc(a, 2 * H, nb * B, 2 * H, W, L.wih1[1], 2 * H, 3 * H, o, st));
    }
    const int nsteps = nf > nb ? nf : nb;
    for (int s = 0; s < nsteps; ++s) {
        GruStep dd[2];
        int n = 0;
        for (int dir = 0; dir < 2; ++dir) {
            const int nd = dir == 0 ? nf : nb;
            if (s >= nd) continue;
            GruStep& x = dd[n++];
            x.d.gi = dir == 0 ? ws.gi1f + (size_t)s * B * 3 * H : ws.gi1b + (size_t)(T - 1 - s - mid) * B * 3 * H;
            x.d.ld_gi = 3 * H;
            x.d.hprev = s > 0 ? ws.h1[dir][(s - 1) & 1] : nullptr; x.d.ld_h = H;
            x.d.whh = W.f + L.whh1[dir]; x.d.bhh = W.f + L.bhh1[dir];
            x.hprev_s = s > 0 ? ws.h1_s[dir][(s - 1) & 1] : NO_SPLIT; x.ld_hs = H;
            x.whh_off = L.whh1[dir];
            if (s == nd - 1) { x.d.hout = g + dir * H; x.d.ld_o = 2 * H; x.d.hs = NO_SPLIT; x.d.ld_s = 0; }
            else { x.d.hout = ws.h1[dir][s & 1]; x.d.ld_o = H; x.d.hs = ws.h1_s[dir][s & 1]; x.d.ld_s = H; }
        }
        RET(gru_step(dd, n, W, B, H, (s > 0 && few.steps-- > 0) ? few.ctas : 0, st));
    }
    return 0;
}

// Side stream + fork/join events so the two encoder streams (pose lifter / GRU image-feature aggregation) run
// concurrently inside one pmce_forward call. One set per (device, caller stream), created on first use (before any graph
// capture: the host driver always runs an eager warm-up first) under a mutex: calls on DIFFERENT streams of one device -
// from one thread or several - never share a side stream or an event, so they are independent as the header promises
// (each call also needs its own workspace: the caller owns that). Two concurrent calls on the SAME stream have no defined
// order among themselves in CUDA either. The sets are the only process-lifetime state besides the layout cache; they are
// never freed (a destroyed-and-recreated stream handle simply reuses its entry).
struct Aux {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, join2 = nullptr, join3 = nullptr;
};
Aux* get_aux(cudaStream_t caller) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Aux> table;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    Aux& a = table[std::make_pair(dev, caller)];           // std::map nodes are address-stable
    if (!a.join2) {
        if (!a.side) {       // highest priority: its few-CTA kernels take the first SMs any lifter kernel frees
            int lo = 0, hi = 0;
            if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
            // PMCE_SIDE_PRIO=0: default priority (A/B knob, read when the side stream of a caller stream is first made)
            if (cudaStreamCreateWithPriority(&a.side, cudaStreamNonBlocking, pmce_env_int("PMCE_SIDE_PRIO", 1) ? hi : lo) != cudaSuccess) return nullptr;
        }
        if (!a.fork && cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (!a.join && cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (!a.fork2 && cudaEventCreateWithFlags(&a.fork2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (!a.join3 && cudaEventCreateWithFlags(&a.join3, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&a.join2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &a;
}

// ---------------------------------------------------------------------------------------------------
// a5 AdaLN gamma/beta, a6-a8 co-evolution block
// ---------------------------------------------------------------------------------------------------
int adaln_gammabeta(const Layout& L, const Weights& W, const float* g, int B, float* gb, const SplitOut& g_s, cudaStream_t st) {
    const int N = PMCE_ADALN_SLOTS * 2 * L.d.coevo_dim, F = L.d.feat_dim;
    RET(split_rows(g, B, F, F, false, g_s, F, st));
    EpiOpt o; o.bias = W.f + L.adaln_b; o.out = gb; o.ld_out = N;
    return linear_tc(g_s, F, B, F, W, L.adaln_w, F, N, o, st);
}

bool mlp_fused_enabled() {
    static int on = -1;      // PMCE_MLP_FUSED=0: keep AdaLN-apply + fc1 GEMM + fc2 GEMM as separate launches (A/B profiling, tests)
    if (on < 0) on = pmce_env_int("PMCE_MLP_FUSED", 1) ? 1 : 0;
    return on == 1;
}

// How the fused Mlp kernel closes the block (mlp_fused.cuh): nothing extra, the feature -> coordinate projection + residual
// (CoevoDecoder.py:189), or the NEXT AdaLayerNorm as the split-bf16 operand of the next projection.
struct MlpTail {
    int epi = MLP_EPI_X;
    size_t f2cw = 0, f2cb = 0; const float* coords_in = nullptr; float* coords_out = nullptr;     // MLP_EPI_F2C
    int slot_next = 0; SplitOut t{nullptr, nullptr};                                              // MLP_EPI_T
};

// x += fc2(gelu(fc1(AdaLN_s2(x)))) over x [B*ntok, 64]   (CoevoDecoder.py:86 / :104)
int adaln_mlp(const Weights& W, int s2, size_t fc1w, size_t fc1b, size_t fc2w, size_t fc2b, float* x, const SplitOut& tmp, const SplitOut& hid,
              const float* gb, int B, int ntok, const MlpTail& tail, cudaStream_t st) {
    const int M = B * ntok;
    if (mlp_fused_enabled()) {
        MlpFusedArgs a;
        memset(&a, 0, sizeof(a));
        a.x = x; a.gb = gb; a.gb_ld = PMCE_ADALN_SLOTS * 128; a.slot = s2; a.slot_next = tail.slot_next;
        a.b1 = W.f + fc1b; a.b2 = W.f + fc2b; a.rows = M; a.ntok = ntok; a.eps = 1e-6f;
        if (tail.epi == MLP_EPI_F2C) { a.wc = W.f + tail.f2cw; a.bc = W.f + tail.f2cb; a.coords_in = tail.coords_in; a.coords_out = tail.coords_out; }
        MlpWeights w{W.hi + fc1w, W.lo + fc1w, W.hi + fc2w, W.lo + fc2w};
        count_launch();
        const int rc = launch_mlp64_fused(tail.epi, w, tail.t, a, st);
        if (rc) { pmce_set_error("mlp64_fused launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
        return 0;
    }
    RET(adaln(x, B, ntok, gb, s2, tmp, st));
    { EpiOpt o; o.bias = W.f + fc1b; o.act = 1; o.outs = hid; o.ld_split = 256; RET(linear_tc(tmp, 64, M, 64, W, fc1w, 64, 256, o, st)); }
    { EpiOpt o; o.bias = W.f + fc2b; o.resid = x; o.ld_resid = 64; o.out = x; o.ld_out = 64; RET(linear_tc(hid, 256, M, 256, W, fc2w, 256, 64, o, st)); }
    if (tail.epi == MLP_EPI_F2C) {
        PLAUNCH(feat2coor_kernel, cdiv(M, 8), 256, 0, st, x, M, W.f + tail.f2cw, W.f + tail.f2cb, tail.coords_in, tail.coords_out);
    } else if (tail.epi == MLP_EPI_T) {
        RET(adaln(x, B, ntok, gb, tail.slot_next, tail.t, st));
    }
    return 0;
}

// residual-stream tail shared by CrossAttentionBlock and Block: x += proj(att); x += fc2(gelu(fc1(AdaLN_2(x))))
int attn_tail(const Weights& W, size_t wp, size_t bp, int s2, size_t fc1w, size_t fc1b, size_t fc2w, size_t fc2b, float* x, const SplitOut& att,
              const SplitOut& tmp, const SplitOut& hid, const float* gb, int B, int ntok, cudaStream_t st, const MlpTail& tail = MlpTail()) {
    const int M = B * ntok;
    { EpiOpt o; o.bias = W.f + bp; o.resid = x; o.ld_resid = 64; o.out = x; o.ld_out = 64; RET(linear_tc(att, 64, M, 64, W, wp, 64, 64, o, st)); }
    return adaln_mlp(W, s2, fc1w, fc1b, fc2w, fc2b, x, tmp, hid, gb, B, ntok, tail, st);
}

int proj64(const SplitOut& a, int M, const Weights& W, size_t w, size_t b, float* out, cudaStream_t st, const float* rowadd = nullptr, int period = 1,
           int N = 64) {
    EpiOpt o; o.bias = W.f + b; o.out = out; o.ld_out = N; o.rowadd = rowadd; o.period = period;
    return linear_tc(a, 64, M, 64, W, w, 64, N, o, st);
}

// Scratch of one attention block over a query stream of n1 rows and a key/value stream of n2 rows (carved from the workspace)
struct AttnScratch {
    SplitOut tq, tk, tv;     // AdaLN outputs (split): [n1,64], [n2,64], [n2,64]   (self-attention uses tq only)
    float *Q, *K, *V;        // projected q/k/v fp32 [n1,64], [n2,64], [n2,64]    (self-attention: Q = qkv [n1,192])
    SplitOut att, hid;       // attention output [n1,64], MLP hidden [n1,256] (split)
    bf16* att_rec;           // [n1, 512] operand records of attn_rows_tc_kernel (vertex stream only; aliases the fp32 qkv buffer)
    CaFolded fold;           // folded per-clip operands (fused vertex cross-attention only)
};

bool ca_fused_enabled() {
    static int on = -1;      // PMCE_CA_FUSED=0: keep the unfused launch sequence (A/B profiling, tests)
    if (on < 0) on = pmce_env_int("PMCE_CA_FUSED", 1) ? 1 : 0;
    return on == 1;
}

int mha_core(int heads, const float* Q, int ldq, const float* K, const float* V, int ldkv, const SplitOut& att, int B, int N1, int N2, cudaStream_t st) {
    const int D = 64 / heads;
    if (D == 32) return flash_attn32(Q, addr_plain(N1, ldq), K, V, addr_plain(N2, ldkv), att, addr_plain(N1, 64), B, heads, N1, N2, st);
    return launch_attn(D, Q, addr_plain(N1, ldq), K, V, addr_plain(N2, ldkv), nullptr, att, addr_plain(N1, 64), B, heads, N1, N2, st);
}

bool ca_fused_ok(int heads, int N1, int N2) { return N1 >= 128 && ca_fused_enabled() && ca_vertex_fused_supported(heads, N2); }

// keys / values of a CrossAttentionBlock: K = Wk AdaLN_k(xk) + bk, V = Wv AdaLN_v(xv) + bv   (xk, xv [B,N2,64])
int cross_attn_kv(const Weights& W, const CaW& w, const float* xk, const float* xv, int N2, const float* gb, int B, const AttnScratch& s, cudaStream_t st) {
    const int n2 = B * N2;
    RET(adaln(xk, B, N2, gb, w.sk, s.tk, st));
    RET(proj64(s.tk, n2, W, w.wk, w.bk, s.K, st));
    RET(adaln(xv, B, N2, gb, w.sv, s.tv, st));
    return proj64(s.tv, n2, W, w.wv, w.bv, s.V, st);
}

// per-clip operands of the fused vertex cross-attention (ca_fused.cuh). joints != nullptr: the whole joint side of coevoblock
// `cw` from the joint coordinates (embed, key projection, AdaLN_k/v, Wk/Wv, fold); else fold given K / V [B,J,64].
JointFoldArgs ca_fold_args(const Weights& W, const CoevoW* cw, const CaW& w, const float* joints, const float* K_in, const float* V_in, float* xq_out,
                           const float* gb, int J, int heads, const CaFolded& f) {
    JointFoldArgs a;
    memset(&a, 0, sizeof(a));
    a.joints = joints; a.K_in = K_in; a.V_in = V_in;
    if (joints) {
        a.wjp = W.f + cw->jprojw; a.bjp = W.f + cw->jprojb; a.jpos = W.f + cw->jpos; a.jq = xq_out ? W.f + cw->jQ : nullptr;
        a.wj2v = W.f + cw->j2vw; a.bj2v = W.f + cw->j2vb; a.j2vk = W.f + cw->j2vK;
        a.wk = W.f + w.wk; a.bk = W.f + w.bk; a.wv = W.f + w.wv; a.bv = W.f + w.bv;
        a.xq_out = xq_out;
    }
    a.wq = W.f + w.wq; a.bq = W.f + w.bq; a.wp = W.f + w.wp; a.bp = W.f + w.bp;
    a.gb = gb; a.gb_ld = PMCE_ADALN_SLOTS * 128; a.slot_k = w.sk; a.slot_v = w.sv; a.slot_q = w.sq;
    a.f = f; a.J = J; a.eps = 1e-6f; a.scale = 1.0f / sqrtf(64.0f / heads);
    return a;
}

int ca_fold_launch(const JointFoldArgs3& args, int nblk, int B, cudaStream_t st) {
    if (!pmce_configure_smem<ca_joint_fold_kernel>(JKV_SMEM)) { pmce_set_error("ca_fold: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); return 10; }
    PLAUNCH(ca_joint_fold_kernel, dim3(B, nblk), JKV_THREADS, JKV_SMEM, st, args);
    return 0;
}

int ca_fold(const Weights& W, const CoevoW* cw, const CaW& w, const float* joints, const float* K_in, const float* V_in, float* xq_out,
            const float* gb, int B, int J, int heads, const CaFolded& f, cudaStream_t st) {
    JointFoldArgs3 args;
    memset(&args, 0, sizeof(args));
    args.blk[0] = ca_fold_args(W, cw, w, joints, K_in, V_in, xq_out, gb, J, heads, f);
    return ca_fold_launch(args, 1, B, st);
}

// xq += proj(MHA(...)) in one pass (ca_fused.cuh); AdaLN_q's gamma/beta, the output bias and log2e are inside the folded operands
int ca_fused_launch(float* xq, int N1, int N2, int B, const CaFolded& f, cudaStream_t st, const CaEmbed* embed = nullptr) {
    CaFusedArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.N1 = N1; a.N2 = N2; a.eps = 1e-6f;
    count_launch();
    const int rc = launch_ca_vertex_fused(xq, f, a, st, embed);
    if (rc) { pmce_set_error("ca_vertex_fused launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

// the query side of a CrossAttentionBlock: xq [B,N1,64] updated in place. Needs projected K / V in s.K / s.V, or (folded) the
// folded operands in s.fold when the fused kernel applies.
int cross_attn_query(const Weights& W, const CaW& w, int heads, float* xq, int N1, int N2, const float* gb, int B, const AttnScratch& s, bool folded,
                     cudaStream_t st, const MlpTail& tail = MlpTail(), const CaEmbed* embed = nullptr) {
    const int n1 = B * N1;
    if (ca_fused_ok(heads, N1, N2)) {
        if (!folded) RET(ca_fold(W, nullptr, w, nullptr, s.K, s.V, nullptr, gb, B, N2, heads, s.fold, st));
        // one pass over the query stream: AdaLN_q, scores, softmax, P V Wp, residual (ca_fused.cuh); then AdaLN_2 + Mlp (mlp_fused.cuh)
        RET(ca_fused_launch(xq, N1, N2, B, s.fold, st, embed));
        return adaln_mlp(W, w.s2, w.fc1w, w.fc1b, w.fc2w, w.fc2b, xq, s.tq, s.hid, gb, B, N1, tail, st);
    }
    RET(adaln(xq, B, N1, gb, w.sq, s.tq, st));
    RET(proj64(s.tq, n1, W, w.wq, w.bq, s.Q, st));
    RET(mha_core(heads, s.Q, 64, s.K, s.V, 64, s.att, B, N1, N2, st));
    return attn_tail(W, w.wp, w.bp, w.s2, w.fc1w, w.fc1b, w.fc2w, w.fc2b, xq, s.att, s.tq, s.hid, gb, B, N1, st, tail);
}

// a6 CrossAttentionBlock.forward (CoevoDecoder.py:82-87): xq [B,N1,64] updated in place; xk, xv [B,N2,64]
int cross_attn_block(const Weights& W, const CaW& w, int heads, float* xq, int N1, const float* xk, const float* xv, int N2, const float* gb, int B,
                     const AttnScratch& s, cudaStream_t st, const MlpTail& tail = MlpTail()) {
    RET(cross_attn_kv(W, w, xk, xv, N2, gb, B, s, st));
    return cross_attn_query(W, w, heads, xq, N1, N2, gb, B, s, false, st, tail);
}

// a7 Block.forward (CoevoDecoder.py:102-105): x [B,N,64] updated in place. t_ready: s.tq already holds AdaLN_1(x) (split), written
// by the previous block's fused Mlp epilogue. tail: how the block's Mlp kernel closes (e.g. the feature -> coordinate projection).
int self_attn_block(const Weights& W, const SaW& w, int heads, float* x, int N, const float* gb, int B, const AttnScratch& s, cudaStream_t st,
                    bool t_ready = false, const MlpTail& tail = MlpTail()) {
    const int n = B * N;
    if (!t_ready) RET(adaln(x, B, N, gb, w.s1, s.tq, st));
    static int rows_on = -1;      // PMCE_ATTN_ROWS=0: fp32 qkv + the chunked online-softmax kernel (A/B profiling; also the path for N > 448)
    if (rows_on < 0) rows_on = pmce_env_int("PMCE_ATTN_ROWS", 1) ? 1 : 0;
    if (rows_on && s.att_rec && attn_rows_tc_supported(heads, N)) {
        // the qkv projection writes the attention kernel's operand tiles (TC_ATTN32 epilogue); whole score rows in TMEM (attn_rows_tc.cuh)
        EpiOpt o; o.bias = W.f + w.qkvb; o.att = s.att_rec; o.qscale = attn_rows_qscale();
        RET(linear_tc(s.tq, 64, n, 64, W, w.qkvw, 64, 192, o, st));
        count_launch();
        const int rc = launch_attn_rows_tc(s.att_rec, s.att, addr_plain(N, 64), B, heads, N, st);
        if (rc) { pmce_set_error("attn_rows_tc launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    } else {
        RET(proj64(s.tq, n, W, w.qkvw, w.qkvb, s.Q, st, nullptr, 1, 192));
        RET(mha_core(heads, s.Q, 192, s.Q + 64, s.Q + 128, 192, s.att, B, N, N, st));
    }
    return attn_tail(W, w.wp, w.bp, w.s2, w.fc1w, w.fc1b, w.fc2w, w.fc2b, x, s.att, s.tq, s.hid, gb, B, N, st, tail);
}

AttnScratch vertex_scratch(const Workspace& ws) {   // query stream = the 431 vertices
    AttnScratch s;
    s.tq = ws.tA_s; s.tk = ws.tJ_s; s.tv = ws.tJ_s; s.Q = ws.qkv_d; s.K = ws.Kj; s.V = ws.Vj; s.att = ws.att_ds; s.hid = ws.hid_ds;
    s.fold = ws.fold[0];
    s.att_rec = reinterpret_cast<bf16*>(ws.qkv_d);       // nv * 512 bf16 = nv * 1024 B <= the fp32 qkv buffer's nv * 192 * 4 B... (see carve)
    return s;
}
AttnScratch joint_scratch(const Workspace& ws) {    // query stream = the J joints
    AttnScratch s;
    s.tq = ws.tJ2_s; s.tk = ws.tB_s; s.tv = ws.tA2_s; s.Q = ws.qkvj; s.K = ws.Kv; s.V = ws.Vv; s.att = ws.attj_s; s.hid = ws.hidj_s;
    memset(&s.fold, 0, sizeof(s.fold));
    s.att_rec = nullptr;
    return s;
}
constexpr int JOINT_HEADS = 8, VERTX_HEADS = 2;      // CoevoDecoder.py:139-140

bool coevo_fused(const pmce_dims_t& d) { return ca_fused_ok(VERTX_HEADS, d.num_vert_ds, d.num_joint) && d.num_joint <= JKV_ROWS; }

// E tables of the embed-mode cross-attention kernel for blocks [k0, k0 + n): weights only, so pmce_forward makes them on the
// image-feature stream
int ca_embed_tables(const Layout& L, const Weights& W, int k0, int n, const Workspace& ws, cudaStream_t st) {
    CaEmbedTableArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n; ++i) {
        const CoevoW& w = L.blk[k0 + i];
        a.bias[i] = W.f + w.vprojb; a.pos[i] = W.f + w.vpos; a.q[i] = W.f + w.vQ; a.out[i] = ws.etab[k0 + i];
    }
    a.n = L.d.num_vert_ds;
    ca_embed_table_kernel<<<dim3(cdiv(a.n * 64, 256), n), 256, 0, st>>>(a);
    CKL();
    return 0;
}

// PMCE_CA_EMBED=1: the vertex cross-attention kernel builds its query stream from the coordinates (embed mode, ca_fused.cuh) instead
// of reading what a separate coevo_embed_kernel launch wrote. Parity-green, one launch and 110 KB/clip of HBM reads less - and
// SLOWER (same box: kernel 22.9 -> 24.1 us at 256 clips, 70.4 -> 78.5 us at 1024; forward 24.7k -> 24.5k clips/s): the kernel is
// bound by the instructions and latencies of its per-item chain, not by DRAM bytes, and the embedding adds ~350 of them per row.
// Off by default; the measurement is the point.
bool ca_embed_enabled() {
    static int on = -1;
    if (on < 0) on = pmce_env_int("PMCE_CA_EMBED", 0) ? 1 : 0;
    return on == 1;
}

// prefolded: the folded operands of this block (ws.fold[k]) and, for block 3, the joint query stream ws.xqj were already made
// by decoder_fold_all (the joint side of all three blocks depends only on the lifter's joints and on gb).
int coevo_block(const Layout& L, const Weights& W, int k, const float* joints, const float* verts_in, const float* gb, int B,
                float* joints_out, float* verts_out, const Workspace& ws, cudaStream_t st, Aux* aux = nullptr, bool prefolded = false) {
    NvtxRange nvtx(k == 0 ? "pmce/coevoblock1" : (k == 1 ? "pmce/coevoblock2" : "pmce/coevoblock3"));
    const pmce_dims_t& d = L.d;
    const CoevoW& w = L.blk[k];
    const int J = d.num_joint, Vd = d.num_vert_ds;
    const int nj = B * J, nv = B * Vd;
    const bool ja = joints_out != nullptr;
    if (ja && !w.joint_alive) { pmce_set_error("coevoblock%d: joint-branch weights are not stored (output is discarded by the reference)", k + 1); return 4; }

    AttnScratch sv = vertex_scratch(ws);
    const bool fused = coevo_fused(d);
    if (prefolded) sv.fold = ws.fold[k];
    // coordinate -> feature (+pos), query streams (+Q embed) (CoevoDecoder.py:177-183); keys proj_j2v(Jf) + j2v_K and
    // proj_v2j(Vf) + v2j_K from the PRE-update features (:183-184)
    if (fused) {
        // the whole joint side of the vertex cross-attention (embed, key projection, AdaLN_k/v, Wk/Wv, fold) per clip in one kernel
        if (!prefolded) RET(ca_fold(W, &w, w.vca, joints, nullptr, nullptr, ja ? ws.xqj : nullptr, gb, B, J, VERTX_HEADS, sv.fold, st));
    } else {
        PLAUNCH(coevo_embed_kernel, cdiv((long long)nj * 16, 256), 256, 0, st, joints, nj, J, W.f + w.jprojw, W.f + w.jprojb, W.f + w.jpos,
                ja ? W.f + w.jQ : nullptr, ws.Jf, ws.Jf_s, ja ? ws.xqj : nullptr);
        RET(proj64(ws.Jf_s, nj, W, w.j2vw, w.j2vb, ws.xkj, st, W.f + w.j2vK, J));
    }
    // vertex query stream: in embed mode the cross-attention kernel builds it from the coordinates itself (ca_fused.cuh)
    const bool emb = fused && ca_embed_enabled();
    if (emb && !prefolded) RET(ca_embed_tables(L, W, k, 1, ws, st));
    if (!emb) {
        PLAUNCH(coevo_embed_kernel, cdiv((long long)nv * 16, 256), 256, 0, st, verts_in, nv, Vd, W.f + w.vprojw, W.f + w.vprojb, W.f + w.vpos, W.f + w.vQ,
                ja ? ws.Vf : nullptr, ja ? ws.Vf_s : NO_SPLIT, ws.xqv);
    }

    if (ja) {
        // the joint branch only reads pre-update features: it runs on the side stream next to the vertex branch
        cudaStream_t sj = st;
        if (aux) {
            CK(cudaEventRecord(aux->fork2, sj));
            CK(cudaStreamWaitEvent(aux->side, aux->fork2, 0));
            sj = aux->side;
        }
        if (emb) {      // the vertex FEATURES (keys / values of the joint cross-attention) are only needed here
            PLAUNCH(coevo_embed_kernel, cdiv((long long)nv * 16, 256), 256, 0, sj, verts_in, nv, Vd, W.f + w.vprojw, W.f + w.vprojb, W.f + w.vpos, nullptr,
                    ws.Vf, ws.Vf_s, nullptr);
        }
        // keys of the joint cross-attention: proj_v2j(Vf) + v2j_K. Only the joint branch reads them, so the projection (a
        // row-embedding epilogue: the direct-store path, 38 us at 64 clips) stays off the vertex branch's stream
        RET(proj64(ws.Vf_s, nv, W, w.v2jw, w.v2jb, ws.xkv, sj, W.f + w.v2jK, Vd));
        const AttnScratch s = joint_scratch(ws);
        // joint cross-attention block: q = joints (J), k/v = vertices (431); 8 heads x 8; then joint self-attention
        MlpTail tj_ca, tj_sa;                                  // the Mlp kernels also produce the next AdaLN / the coordinates
        tj_ca.epi = MLP_EPI_T; tj_ca.slot_next = w.jsa.s1; tj_ca.t = s.tq;
        tj_sa.epi = MLP_EPI_F2C; tj_sa.f2cw = w.jf2cw; tj_sa.f2cb = w.jf2cb; tj_sa.coords_in = joints; tj_sa.coords_out = joints_out;
        RET(cross_attn_block(W, w.jca, JOINT_HEADS, ws.xqj, J, ws.xkv, ws.Vf, Vd, gb, B, s, sj, tj_ca));
        RET(self_attn_block(W, w.jsa, JOINT_HEADS, ws.xqj, J, gb, B, s, sj, true, tj_sa));
        if (aux) CK(cudaEventRecord(aux->join2, aux->side));
    }

    // vertex cross-attention block: q = vertices (431), k/v = joints (J); 2 heads x 32; then vertex self-attention 431 x 431
    if (!fused) RET(cross_attn_kv(W, w.vca, ws.xkj, ws.Jf, J, gb, B, sv, st));
    MlpTail tv_ca, tv_sa;
    tv_ca.epi = MLP_EPI_T; tv_ca.slot_next = w.vsa.s1; tv_ca.t = sv.tq;
    tv_sa.epi = MLP_EPI_F2C; tv_sa.f2cw = w.vf2cw; tv_sa.f2cb = w.vf2cb; tv_sa.coords_in = verts_in; tv_sa.coords_out = verts_out;
    CaEmbed ce{verts_in, ws.etab[k], W.f + w.vprojw};
    RET(cross_attn_query(W, w.vca, VERTX_HEADS, ws.xqv, Vd, J, gb, B, sv, fused, st, tv_ca, emb ? &ce : nullptr));
    RET(self_attn_block(W, w.vsa, VERTX_HEADS, ws.xqv, Vd, gb, B, sv, st, true, tv_sa));
    // with a side stream the CALLER joins (aux->join2): decoder_back puts the up-sampling before the join
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a9 tail: upsample_conv + linear_cur residual
// ---------------------------------------------------------------------------------------------------
// linear_cur1-3(relu(y_mid)) laid out like the mesh ([B,6890,3]); depends only on the GRU output, so pmce_forward runs it
// on the image-feature stream while the pose stream is still busy (169 MB of weights streamed once per batch)
int mesh_residual(const Layout& L, const Weights& W, const float* g, int B, const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int V = d.num_vert, F = d.feat_dim;
    RET(split_rows(g, B, F, F, true, ws.gr_s, F, st));     // relu(y_mid)
    EpiOpt o; o.bias = W.f + L.lc_b; o.out = ws.lc_mesh; o.mapped = true;
    o.rmap.div = 1; o.rmap.s0 = (long long)V * 3; o.rmap.s1 = 0;
    o.cmap.div = V; o.cmap.s0 = 1; o.cmap.s1 = 3;
    return linear_tc(ws.gr_s, F, B, F, W, L.lc_w, F, 3 * V, o, st);
}

// mesh[b,o,l] = b_up[o] + sum_{c,k} W_up[o,c,k] verts3[b,c,l+k-1] + lc_mesh[b,o,l]
int mesh_upsample(const Layout& L, const Weights& W, const float* verts3, int B, float* mesh, const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int Vd = d.num_vert_ds, V = d.num_vert, ldk = L.ups_ld;
    PLAUNCH(upsample_im2col_kernel, cdiv((long long)B * 3 * ldk, 256), 256, 0, st, verts3, B, Vd, ldk, nullptr, ws.im2col_s);
    EpiOpt o; o.bias = W.f + L.ups_b; o.out = mesh; o.resid = ws.lc_mesh; o.mapped = true;
    o.force_bn = pmce_env_int("PMCE_UPS_BN", 0);     // A/B of the tile width of this skinny GEMM (M = 3 B rows, N = 6890; read live)
    o.rmap.div = 3; o.rmap.s0 = (long long)V * 3; o.rmap.s1 = 1;
    o.cmap.div = 1; o.cmap.s0 = 3; o.cmap.s1 = 0;
    return linear_tc(ws.im2col_s, ldk, B * 3, ldk, W, L.ups_w, ldk, V, o, st);
}

int mesh_epilogue(const Layout& L, const Weights& W, const float* verts3, const float* g, int B, float* mesh, const Workspace& ws, cudaStream_t st) {
    RET(mesh_residual(L, W, g, B, ws, st));
    return mesh_upsample(L, W, verts3, B, mesh, ws, st);
}

int prepare_feat(const Layout& L, const float* img_feat, int nfr, const Workspace& ws, cudaStream_t st) {
    const int F = L.d.feat_dim;
    return split_rows(img_feat, nfr, F, F, false, ws.feat_s, F, st);
}

// image-feature stream of the two-stream encoder: GRU -> y[T//2] -> all AdaLN gamma/beta (independent of the pose stream)
// gb_ready: recorded once gamma/beta exist - all the co-evolution blocks wait for; the linear_cur mesh residual behind it (169 MB of
// weights streamed once, needed only by the final up-sampling) then runs beside the decoder's latency-bound kernels
int decoder_front(const Layout& L, const Weights& W, int B, int nfr, int fstride, const Workspace& ws, cudaStream_t st, GruFew few = GruFew{0, 0},
                  cudaEvent_t gb_ready = nullptr) {
    NvtxRange nvtx("pmce/image-feature stream (GRU + AdaLN gamma/beta + linear_cur)");
    RET(gru_mid(L, W, B, nfr, fstride, ws.g, ws, st, few));
    RET(adaln_gammabeta(L, W, ws.g, B, ws.gb, ws.g_s, st));
    if (gb_ready) CK(cudaEventRecord(gb_ready, st));
    if (coevo_fused(L.d) && ca_embed_enabled()) RET(ca_embed_tables(L, W, 0, 3, ws, st));
    return mesh_residual(L, W, ws.g, B, ws, st);
}

int decoder_back(const Layout& L, const Weights& W, const float* joints, const int32_t* vj, int B, float* cam_pose, float* cam_mesh,
                 float* verts0_out, const Workspace& ws, cudaStream_t st, Aux* aux = nullptr) {
    NvtxRange nvtx("pmce/decoder (3 co-evolution blocks + up-sampling)");
    const pmce_dims_t& d = L.d;
    const int J = d.num_joint, Vd = d.num_vert_ds;
    float* v0 = verts0_out ? verts0_out : ws.verts[2];
    PLAUNCH(gather_verts_kernel, cdiv((long long)B * Vd * 3, 256), 256, 0, st, joints, vj, B, J, Vd, v0);
    const bool pre = coevo_fused(d);
    if (pre) {   // every block reads the SAME joints (CoevoDecoder.py:235-237 pass P, not P^k): one launch makes all three joint sides
        JointFoldArgs3 args;
        memset(&args, 0, sizeof(args));
        for (int k = 0; k < 3; ++k)
            args.blk[k] = ca_fold_args(W, &L.blk[k], L.blk[k].vca, joints, nullptr, nullptr, k == 2 ? ws.xqj : nullptr, ws.gb, J, VERTX_HEADS, ws.fold[k]);
        RET(ca_fold_launch(args, 3, B, st));
    }
    RET(coevo_block(L, W, 0, joints, v0, ws.gb, B, nullptr, ws.verts[0], ws, st, nullptr, pre));
    RET(coevo_block(L, W, 1, joints, ws.verts[0], ws.gb, B, nullptr, ws.verts[1], ws, st, nullptr, pre));
    RET(coevo_block(L, W, 2, joints, ws.verts[1], ws.gb, B, cam_pose, ws.verts[0], ws, st, aux, pre));
    if (aux) CK(cudaStreamWaitEvent(st, aux->join3, 0));     // the linear_cur residual (image-feature stream) feeds the up-sampling's epilogue
    RET(mesh_upsample(L, W, ws.verts[0], B, cam_mesh, ws, st));
    if (aux) CK(cudaStreamWaitEvent(st, aux->join2, 0));     // block 3's joint branch (cam_pose) ran beside the vertex branch and the up-sampling
    return 0;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
#define GET_LAYOUT()                                   \
    const Layout* Lp = pmce_get_layout(dims);          \
    if (!Lp) return 1;                                 \
    const Layout& L = *Lp;                             \
    if (!weights) { pmce_set_error("weights is NULL"); return 2; } \
    const Weights W = weights_view(L, weights);        \
    cudaStream_t st = (cudaStream_t)stream;

extern "C" size_t pmce_weights_bytes(const pmce_dims_t* dims) {
    const Layout* L = pmce_get_layout(dims);
    return L ? L->total_floats * 8 : 0;     // fp32 + bf16 hi + bf16 lo
}

extern "C" size_t pmce_workspace_bytes(const pmce_dims_t* dims, int B) {
    const Layout* Lp = pmce_get_layout(dims);
    if (!Lp || B < 1) return 0;
    return carve(*dims, B, nullptr).bytes;
}

extern "C" int pmce_pack_weights(const pmce_dims_t* dims, void* weights, void* stream) {
    const Layout* Lp = pmce_get_layout(dims);
    if (!Lp) return 1;
    if (!weights) { pmce_set_error("weights is NULL"); return 2; }
    const Weights W = weights_view(*Lp, weights);
    const size_t n = Lp->total_floats;     // multiple of 64
    SplitOut o{const_cast<bf16*>(W.hi), const_cast<bf16*>(W.lo)};
    // split the whole fp32 region in one pass, viewed as [n/64, 64]
    return split_rows(W.f, (int)(n / 64), 64, 64, false, o, 64, (cudaStream_t)stream);
}

extern "C" int pmce_lifter_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat, int B,
                                   float* pose3d, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    RET(prepare_feat(L, img_feat, B * dims->seqlen, ws, st));
    return lifter(L, W, pose2d, B, B * dims->seqlen, dims->seqlen, pose3d, ws, st);
}

extern "C" int pmce_gru_mid(const pmce_dims_t* dims, const void* weights, const float* img_feat, int B, float* g, void* workspace,
                            size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    RET(prepare_feat(L, img_feat, B * dims->seqlen, ws, st));
    // standalone: 128-CTA steps unless PMCE_GRU_FEW_STEPS asks for few-CTA ones (tests, tools/stage_times.py)
    GruFew few = gru_few_plan(L.d, B);
    if (pmce_env_int("PMCE_GRU_FEW_STEPS", 0) < 0) few.steps = 0;
    return gru_mid(L, W, B, B * dims->seqlen, dims->seqlen, g, ws, st, few);
}

extern "C" int pmce_adaln_gammabeta(const pmce_dims_t* dims, const void* weights, const float* g, int B, float* gb, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return adaln_gammabeta(L, W, g, B, gb, ws.g_s, st);
}

extern "C" int pmce_coevo_block(const pmce_dims_t* dims, const void* weights, int block, const float* joints, const float* verts_in,
                                const float* gb, int B, float* joints_out, float* verts_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
    GET_LAYOUT();
    if (block < 1 || block > 3) { pmce_set_error("block must be 1..3 (got %d)", block); return 2; }
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return coevo_block(L, W, block - 1, joints, verts_in, gb, B, joints_out, verts_out, ws, st);
}

// which: 0 = the joint stream's block (queries = J joints), 1 = the vertex stream's block (queries = 431 vertices)
static int attn_block_args(const Layout& L, int block, int which, const char* what) {
    if (block < 1 || block > 3) { pmce_set_error("%s: block must be 1..3 (got %d)", what, block); return 2; }
    if (which != 0 && which != 1) { pmce_set_error("%s: which must be 0 (joint stream) or 1 (vertex stream)", what); return 2; }
    if (which == 0 && !L.blk[block - 1].joint_alive) {
        pmce_set_error("%s: coevoblock%d joint-branch weights are not stored (output is discarded by the reference)", what, block);
        return 4;
    }
    return 0;
}

extern "C" int pmce_cross_attn_block(const pmce_dims_t* dims, const void* weights, int block, int which, const float* xq, const float* xk,
                                     const float* xv, const float* gb, int B, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    GET_LAYOUT();
    RET(attn_block_args(L, block, which, "pmce_cross_attn_block"));
    if (!xq || !xk || !xv || !gb || !out) { pmce_set_error("pmce_cross_attn_block: NULL argument"); return 2; }
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    const int J = dims->num_joint, Vd = dims->num_vert_ds;
    const int N1 = which ? Vd : J, N2 = which ? J : Vd;
    if (out != xq) CK(cudaMemcpyAsync(out, xq, (size_t)B * N1 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const CoevoW& w = L.blk[block - 1];
    return which ? cross_attn_block(W, w.vca, VERTX_HEADS, out, N1, xk, xv, N2, gb, B, vertex_scratch(ws), st)
                 : cross_attn_block(W, w.jca, JOINT_HEADS, out, N1, xk, xv, N2, gb, B, joint_scratch(ws), st);
}

extern "C" size_t pmce_ca_fold_bytes(int B) { return B < 1 ? 0 : (size_t)B * (4 * CAF_NS * 64 * 2 + CAF_NS * 4); }

extern "C" int pmce_ca_vertex_fused(const pmce_dims_t* dims, const void* weights, int block, float* xq, const float* K, const float* V,
                                    const float* gb, int B, void* t_hi, void* t_lo, void* fold_ws, int fold, void* stream) {
    GET_LAYOUT();
    if (block < 1 || block > 3) { pmce_set_error("pmce_ca_vertex_fused: block must be 1..3 (got %d)", block); return 2; }
    if (!xq || !K || !V || !gb || !fold_ws || B < 1 || (t_hi == nullptr) != (t_lo == nullptr)) { pmce_set_error("pmce_ca_vertex_fused: bad argument"); return 2; }
    if ((uintptr_t)fold_ws & 255) { pmce_set_error("pmce_ca_vertex_fused: fold_ws must be 256-byte aligned"); return 2; }
    const int J = dims->num_joint, Vd = dims->num_vert_ds;
    if (!ca_vertex_fused_supported(VERTX_HEADS, J) || Vd < 128) { pmce_set_error("pmce_ca_vertex_fused: unsupported shape (J=%d, Vd=%d)", J, Vd); return 3; }
    const CaW& w = L.blk[block - 1].vca;
    CaFolded f;
    f.kq_hi = (bf16*)fold_ws; f.kq_lo = f.kq_hi + (size_t)B * CAF_NS * 64; f.vp_hi = f.kq_lo + (size_t)B * CAF_NS * 64;
    f.vp_lo = f.vp_hi + (size_t)B * CAF_NS * 64; f.sb = (float*)(f.vp_lo + (size_t)B * CAF_NS * 64);
    if (fold) RET(ca_fold(W, nullptr, w, nullptr, K, V, nullptr, gb, B, J, VERTX_HEADS, f, st));
    RET(ca_fused_launch(xq, Vd, J, B, f, st));
    if (t_hi && t_lo) {          // optional: AdaLN_2 of the result, split-bf16 (the A operand of the block's fc1 when the Mlp is not fused)
        SplitOut t{(bf16*)t_hi, (bf16*)t_lo};
        RET(adaln(xq, B, Vd, gb, w.s2, t, st));
    }
    return 0;
}

extern "C" int pmce_ca_vertex_fused_embed(const pmce_dims_t* dims, const void* weights, int block, const float* coords, float* xq_out,
                                          const float* K, const float* V, const float* gb, int B, void* fold_ws, int fold, float* table_ws,
                                          void* stream) {
    GET_LAYOUT();
    if (block < 1 || block > 3) { pmce_set_error("pmce_ca_vertex_fused_embed: block must be 1..3 (got %d)", block); return 2; }
    if (!coords || !xq_out || !K || !V || !gb || !fold_ws || !table_ws || B < 1) { pmce_set_error("pmce_ca_vertex_fused_embed: bad argument"); return 2; }
    if (((uintptr_t)fold_ws & 255) || ((uintptr_t)table_ws & 255)) { pmce_set_error("pmce_ca_vertex_fused_embed: fold_ws / table_ws must be 256-byte aligned"); return 2; }
    const int J = dims->num_joint, Vd = dims->num_vert_ds;
    if (!ca_vertex_fused_supported(VERTX_HEADS, J) || Vd < 128) { pmce_set_error("pmce_ca_vertex_fused_embed: unsupported shape (J=%d, Vd=%d)", J, Vd); return 3; }
    const CoevoW& cw = L.blk[block - 1];
    const CaW& w = cw.vca;
    CaFolded f;
    f.kq_hi = (bf16*)fold_ws; f.kq_lo = f.kq_hi + (size_t)B * CAF_NS * 64; f.vp_hi = f.kq_lo + (size_t)B * CAF_NS * 64;
    f.vp_lo = f.vp_hi + (size_t)B * CAF_NS * 64; f.sb = (float*)(f.vp_lo + (size_t)B * CAF_NS * 64);
    if (fold) {
        RET(ca_fold(W, nullptr, w, nullptr, K, V, nullptr, gb, B, J, VERTX_HEADS, f, st));
        CaEmbedTableArgs ta;
        memset(&ta, 0, sizeof(ta));
        ta.bias[0] = W.f + cw.vprojb; ta.pos[0] = W.f + cw.vpos; ta.q[0] = W.f + cw.vQ; ta.out[0] = table_ws; ta.n = Vd;
        ca_embed_table_kernel<<<dim3(cdiv(Vd * 64, 256), 1), 256, 0, st>>>(ta);
        CKL();
    }
    CaEmbed ce{coords, table_ws, W.f + cw.vprojw};
    return ca_fused_launch(xq_out, Vd, J, B, f, st, &ce);
}

extern "C" int pmce_self_attn_block(const pmce_dims_t* dims, const void* weights, int block, int which, const float* x, const float* gb,
                                    int B, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    RET(attn_block_args(L, block, which, "pmce_self_attn_block"));
    if (!x || !gb || !out) { pmce_set_error("pmce_self_attn_block: NULL argument"); return 2; }
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    const int N = which ? dims->num_vert_ds : dims->num_joint;
    if (out != x) CK(cudaMemcpyAsync(out, x, (size_t)B * N * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const CoevoW& w = L.blk[block - 1];
    return which ? self_attn_block(W, w.vsa, VERTX_HEADS, out, N, gb, B, vertex_scratch(ws), st)
                 : self_attn_block(W, w.jsa, JOINT_HEADS, out, N, gb, B, joint_scratch(ws), st);
}

extern "C" int pmce_mesh_epilogue(const pmce_dims_t* dims, const void* weights, const float* verts3, const float* g, int B,
                                  float* cam_mesh, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return mesh_epilogue(L, W, verts3, g, B, cam_mesh, ws, st);
}

extern "C" int pmce_decoder_forward(const pmce_dims_t* dims, const void* weights, const float* joints, const float* img_feat,
                                    const int32_t* vj_relation, int B, float* cam_pose, float* cam_mesh, float* verts0_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    RET(prepare_feat(L, img_feat, B * dims->seqlen, ws, st));
    RET(decoder_front(L, W, B, B * dims->seqlen, dims->seqlen, ws, st));
    return decoder_back(L, W, joints, vj_relation, B, cam_pose, cam_mesh, verts0_out, ws, st);
}

// whole forward over B windows of a track of nfr frames (window b = frames [b*fstride, b*fstride+T)); fstride == T: a batch of clips
static int forward_windows(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat, const int32_t* vj_relation,
                           int B, int nfr, int fstride, float* cam_mesh, float* cam_pose, float* pose3d, void* workspace, size_t workspace_bytes,
                           void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    RET(prepare_feat(L, img_feat, nfr, ws, st));
    // fork: image-feature stream (GRU + AdaLN gamma/beta) on the side stream, pose stream (lifter) on the caller's stream
    Aux* aux = get_aux(st);
    if (!aux) { pmce_set_error("could not create the side stream/events: %s", cudaGetErrorString(cudaGetLastError())); return 10; }
    CK(cudaEventRecord(aux->fork, st));
    CK(cudaStreamWaitEvent(aux->side, aux->fork, 0));
    const GruFew few = gru_few_plan(L.d, B);
    // PMCE_LC_LATE=0: join only after the linear_cur residual (A/B; read live)
    const bool lc_late = pmce_env_int("PMCE_LC_LATE", 1) != 0 && !(skip_mask() & 2);
    if (!(skip_mask() & 2)) { PdlScope sc(PMCE_PDL_SIDE); RET(decoder_front(L, W, B, nfr, fstride, ws, aux->side, few, lc_late ? aux->join : nullptr)); }
    if (!lc_late) CK(cudaEventRecord(aux->join, aux->side));
    CK(cudaEventRecord(aux->join3, aux->side));
    tc_sm_reserve = few.steps > 0 ? few.ctas : 0;
    int lrc = 0;
    if (!(skip_mask() & 4)) { PdlScope sc(PMCE_PDL_LIFTER); lrc = lifter(L, W, pose2d, B, nfr, fstride, pose3d, ws, st); }
    tc_sm_reserve = 0;
    if (lrc) return lrc;
    CK(cudaStreamWaitEvent(st, aux->join, 0));
    if (skip_mask() & 8) return 0;
    return decoder_back(L, W, ws.joints_m, vj_relation, B, cam_pose, cam_mesh, nullptr, ws, st, aux);
}

extern "C" int pmce_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat,
                            const int32_t* vj_relation, int B, float* cam_mesh, float* cam_pose, float* pose3d, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!dims) { pmce_set_error("dims is NULL"); return 2; }
    return forward_windows(dims, weights, pose2d, img_feat, vj_relation, B, B * dims->seqlen, dims->seqlen, cam_mesh, cam_pose, pose3d, workspace,
                           workspace_bytes, stream);
}

extern "C" int pmce_sliding_windows(const pmce_dims_t* dims, int num_frames, int stride) {
    if (!dims || stride < 1 || num_frames < dims->seqlen) return 0;
    return (num_frames - dims->seqlen) / stride + 1;
}

extern "C" int pmce_forward_sliding(const pmce_dims_t* dims, const void* weights, const float* pose2d_seq, const float* img_feat_seq,
                                    const int32_t* vj_relation, int num_frames, int stride, float* cam_mesh, float* cam_pose, float* pose3d,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!dims) { pmce_set_error("dims is NULL"); return 2; }
    const int nwin = pmce_sliding_windows(dims, num_frames, stride);
    if (nwin < 1) { pmce_set_error("pmce_forward_sliding: need num_frames >= seqlen (%d) and stride >= 1 (got %d frames, stride %d)", dims->seqlen, num_frames, stride); return 2; }
    if (stride > dims->seqlen) { pmce_set_error("pmce_forward_sliding: stride %d > seqlen %d leaves gaps; use pmce_forward on the clips", stride, dims->seqlen); return 2; }
    const int nfr = (nwin - 1) * stride + dims->seqlen;       // frames the windows actually cover
    return forward_windows(dims, weights, pose2d_seq, img_feat_seq, vj_relation, nwin, nfr, stride, cam_mesh, cam_pose, pose3d, workspace,
                           workspace_bytes, stream);
}

extern "C" size_t pmce_io_bytes(const pmce_dims_t* dims, int B) {
    if (!dims || B < 1) return 0;
    const size_t J = dims->num_joint, T = dims->seqlen, F = dims->feat_dim, V = dims->num_vert;
    auto al = [](size_t n) { return (n * 4 + 255) / 256 * 256; };
    return al((size_t)B * T * J * 2) + al((size_t)B * T * F) + al((size_t)B * V * 3) + 2 * al((size_t)B * J * 3);
}

extern "C" int pmce_forward_host(const pmce_dims_t* dims, const void* weights, const float* h_pose2d, const float* h_img_feat,
                                 const int32_t* d_vj_relation, int B, float* h_cam_mesh, float* h_cam_pose, float* h_pose3d,
                                 void* d_io, void* workspace, size_t workspace_bytes, void* stream) {
    if (!dims) { pmce_set_error("dims is NULL"); return 2; }
    if (!d_io) { pmce_set_error("d_io is NULL"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t J = dims->num_joint, T = dims->seqlen, F = dims->feat_dim, V = dims->num_vert;
    Carver c{(char*)d_io, 0};
    float* d_p2d = c.f32((size_t)B * T * J * 2);
    float* d_feat = c.f32((size_t)B * T * F);
    float* d_mesh = c.f32((size_t)B * V * 3);
    float* d_pose = c.f32((size_t)B * J * 3);
    float* d_p3d = c.f32((size_t)B * J * 3);
    CK(cudaMemcpyAsync(d_p2d, h_pose2d, (size_t)B * T * J * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_feat, h_img_feat, (size_t)B * T * F * sizeof(float), cudaMemcpyHostToDevice, st));
    RET(pmce_forward(dims, weights, d_p2d, d_feat, d_vj_relation, B, d_mesh, d_pose, d_p3d, workspace, workspace_bytes, stream));
    CK(cudaMemcpyAsync(h_cam_mesh, d_mesh, (size_t)B * V * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_cam_pose, d_pose, (size_t)B * J * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_pose3d, d_p3d, (size_t)B * J * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int pmce_jregress(const int32_t* row_ptr, const int32_t* cols, const float* vals, int R, const float* mesh, int num_vert,
                             int B, float scale, float* out, void* stream) {
    if (!row_ptr || !cols || !vals || !mesh || !out || R < 1 || B < 1) { pmce_set_error("pmce_jregress: bad argument"); return 2; }
    jregress_kernel<<<cdiv((long long)B * R * 3, 128), 128, 0, (cudaStream_t)stream>>>(row_ptr, cols, vals, R, mesh, num_vert, B, scale, out);
    CKL();
    return 0;
}

extern "C" int pmce_eval_errors(const int32_t* row_ptr, const int32_t* cols, const float* vals, int R, const float* cam_mesh,
                                const float* gt_mesh, const float* gt_pose, const int32_t* eval_joints, int n_eval, int num_vert, int B,
                                float scale, float* pred_pose, float* clip_err, float* mean_err, void* stream) {
    if (!row_ptr || !cols || !vals || !cam_mesh || !gt_mesh || !gt_pose || !eval_joints || !pred_pose || !clip_err || !mean_err) {
        pmce_set_error("pmce_eval_errors: NULL argument"); return 2;
    }
    if (R < 1 || B < 1 || num_vert < 1 || n_eval < 1 || n_eval > 256) { pmce_set_error("pmce_eval_errors: bad size (R=%d B=%d V=%d n_eval=%d)", R, B, num_vert, n_eval); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    jregress_kernel<<<cdiv((long long)B * R * 3, 128), 128, 0, st>>>(row_ptr, cols, vals, R, cam_mesh, num_vert, B, scale, pred_pose);
    CKL();
    if (((uintptr_t)cam_mesh | (uintptr_t)gt_mesh) & 7) { pmce_set_error("pmce_eval_errors: meshes must be 8-byte aligned"); return 2; }
    if ((num_vert & 1) && B > 1) { pmce_set_error("pmce_eval_errors: odd vertex counts are supported for B = 1 only (8-byte loads)"); return 2; }
    eval_err_kernel<<<B, EVAL_THREADS, 0, st>>>(cam_mesh, gt_mesh, pred_pose, gt_pose, eval_joints, n_eval, R, num_vert, scale, clip_err);
    CKL();
    eval_mean_kernel<<<1, 64, 0, st>>>(clip_err, B, mean_err);
    CKL();
    return 0;
}

extern "C" int pmce_linear(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                           void* stream) {
    if (!x || !weight || !out || M < 1 || N < 1 || K < 4 || (K & 3)) { pmce_set_error("pmce_linear: bad argument (K must be a multiple of 4)"); return 2; }
    GemmEpi e = gemm_epi_plain(N);
    e.bias = bias; e.act = act;
    CKG(launch_gemm_tn(x, K, weight, K, out, M, N, K, e, (cudaStream_t)stream));
    return 0;
}

extern "C" size_t pmce_linear_tc_scratch_bytes(int M, int N, int K) {
    auto al = [](size_t n) { return (n * 2 + 255) / 256 * 256; };
    return 2 * (al((size_t)M * K) + al((size_t)N * K));
}

extern "C" int pmce_linear_tc(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                              void* scratch, size_t scratch_bytes, void* stream) {
    if (!x || !weight || !out || !scratch || M < 1 || N < 1 || K < 8 || (K & 7)) { pmce_set_error("pmce_linear_tc: bad argument (K must be a multiple of 8)"); return 2; }
    if (scratch_bytes < pmce_linear_tc_scratch_bytes(M, N, K)) { pmce_set_error("pmce_linear_tc: scratch too small"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    Carver c{(char*)scratch, 0};
    SplitOut a = c.split((size_t)M * K), w = c.split((size_t)N * K);
    RET(split_rows(x, M, K, K, false, a, K, st));
    RET(split_rows(weight, N, K, K, false, w, K, st));
    TcOperand A{a.hi, a.lo, M, K, K}, Wm{w.hi, w.lo, N, K, K};
    TcEpi e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.act = act; e.out_f32 = out; e.ld_out = N; e.rowadd_period = 1;
    count_launch();
    const int rc = launch_linear_tc(A, Wm, e, st);
    if (rc) { pmce_set_error("pmce_linear_tc: launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

extern "C" int pmce_split_bf16(const float* x, int rows, int cols, void* hi, void* lo, void* stream) {
    if (!x || !hi || !lo || rows < 1 || cols < 4 || (cols & 3)) { pmce_set_error("pmce_split_bf16: bad argument"); return 2; }
    SplitOut o{(bf16*)hi, (bf16*)lo};
    return split_rows(x, rows, cols, cols, false, o, cols, (cudaStream_t)stream);
}

extern "C" int pmce_linear_tc_presplit(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int M,
                                       int N, int K, int act, float* out, void* out_hi, void* out_lo, const float* resid, void* stream) {
    if (!x_hi || !x_lo || !w_hi || !w_lo || (!out && !out_hi) || (out_hi && !out_lo) || M < 1 || N < 1 || K < 8 || (K & 7)) {
        pmce_set_error("pmce_linear_tc_presplit: bad argument"); return 2;
    }
    TcOperand A{(const bf16*)x_hi, (const bf16*)x_lo, M, K, K}, Wm{(const bf16*)w_hi, (const bf16*)w_lo, N, K, K};
    TcEpi e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.act = act; e.out_f32 = out; e.ld_out = N; e.rowadd_period = 1;
    e.out_hi = (bf16*)out_hi; e.out_lo = (bf16*)out_lo; e.ld_split = N; e.resid = resid; e.ld_resid = N;
    count_launch();
    const int rc = launch_linear_tc(A, Wm, e, (cudaStream_t)stream);
    if (rc) { pmce_set_error("pmce_linear_tc_presplit: launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

// ---- SMPL LBS --------------------------------------------------------------------------------------
#define SMPL_V 6890
#define SMPL_LDK 224
extern "C" int smpl_blend_ld(void) { return SMPL_LDK; }
#define SMPL_NPAD 20672      /* 6890*3 rounded up to the GEMM epilogue's 16-column chunk */
extern "C" size_t smpl_workspace_bytes(int B) {
    if (B < 1) return 0;
    auto al = [](size_t n) { return (n + 255) / 256 * 256; };
    return al((size_t)B * SMPL_LDK * 4) + al((size_t)B * 288 * 4) + al((size_t)B * SMPL_NPAD * 4) + 2 * al((size_t)B * SMPL_LDK * 2);
}

extern "C" int smpl_lbs_forward_sparse(const float* blend, const void* blend_hi, const void* blend_lo, const float* v_template,
                                       const float* j_template, const float* j_shapedirs, const float* skin_weights, const int32_t* skin_idx4,
                                       const float* skin_w4, const int32_t* parents, const float* pose, const float* betas, const float* trans,
                                       int B, float out_scale, float* verts, float* joints, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int smpl_lbs_forward_scaled(const float* blend, const void* blend_hi, const void* blend_lo, const float* v_template,
                                       const float* j_template, const float* j_shapedirs, const float* skin_weights, const int32_t* parents,
                                       const float* pose, const float* betas, const float* trans, int B, float out_scale, float* verts,
                                       float* joints, void* workspace, size_t workspace_bytes, void* stream) {
    return smpl_lbs_forward_sparse(blend, blend_hi, blend_lo, v_template, j_template, j_shapedirs, skin_weights, nullptr, nullptr, parents, pose,
                                   betas, trans, B, out_scale, verts, joints, workspace, workspace_bytes, stream);
}

extern "C" int smpl_lbs_forward_sparse(const float* blend, const void* blend_hi, const void* blend_lo, const float* v_template,
                                       const float* j_template, const float* j_shapedirs, const float* skin_weights, const int32_t* skin_idx4,
                                       const float* skin_w4, const int32_t* parents, const float* pose, const float* betas, const float* trans,
                                       int B, float out_scale, float* verts, float* joints, void* workspace, size_t workspace_bytes, void* stream) {
    if ((!blend && !blend_hi) || (blend_hi && !blend_lo) || !v_template || !j_template || !j_shapedirs || (!skin_weights && !skin_idx4) ||
        ((skin_idx4 == nullptr) != (skin_w4 == nullptr)) || !parents || !pose || !betas ||
        !verts || !joints) {
        pmce_set_error("smpl_lbs_forward: NULL argument"); return 2;
    }
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    if (!workspace || workspace_bytes < smpl_workspace_bytes(B)) { pmce_set_error("smpl workspace too small"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    Carver c{(char*)workspace, 0};
    float* coef = c.f32((size_t)B * SMPL_LDK);
    float* Amat = c.f32((size_t)B * 288);
    float* vposed = c.f32((size_t)B * SMPL_NPAD);
    SplitOut coef_s = c.split((size_t)B * SMPL_LDK);
    smpl_pose_kernel<<<B, 32, 0, st>>>(pose, betas, trans, j_template, j_shapedirs, parents, B, SMPL_LDK, coef, Amat, joints, out_scale,
                                       blend_hi ? coef_s : NO_SPLIT);
    CKL();
    // v_posed[b, v*3+c] = v_template + [shapedirs | posedirs] . [betas | pose_map]   (smpl_layer.py:93-99)
    int ld_vp;
    if (blend_hi) {
        // tensor cores (bf16x3, as every projection of the forward): blend_hi/lo and v_template are padded to SMPL_NPAD rows
        Weights Wb{nullptr, (const bf16*)blend_hi, (const bf16*)blend_lo};      // coef_s: written by the pose kernel
        EpiOpt o; o.bias = v_template; o.out = vposed; o.ld_out = SMPL_NPAD;
        RET(linear_tc(coef_s, SMPL_LDK, B, SMPL_LDK, Wb, 0, SMPL_LDK, SMPL_NPAD, o, st));
        ld_vp = SMPL_NPAD;
    } else {
        GemmEpi e = gemm_epi_plain(SMPL_V * 3);      // exact fp32 CUDA-core GEMM
        e.bias = v_template;
        CKG(launch_gemm_tn(coef, SMPL_LDK, blend, SMPL_LDK, vposed, B, SMPL_V * 3, SMPL_LDK, e, st));
        ld_vp = SMPL_V * 3;
    }
    dim3 grid(cdiv(SMPL_V, 256), cdiv(B, SMPL_SKIN_NB));
    if (skin_idx4)      // <= 4 joints per vertex (every shipped SMPL model): the sparse stream kernel, one thread per vertex pair
        smpl_skin4_kernel<<<dim3(cdiv((SMPL_V + 1) / 2, 256), B), 256, 0, st>>>(vposed, Amat, reinterpret_cast<const int4*>(skin_idx4), reinterpret_cast<const float4*>(skin_w4), trans,
                                                SMPL_V, B, ld_vp, verts, out_scale);
    else
        smpl_skin_kernel<<<grid, 256, 0, st>>>(vposed, Amat, skin_weights, trans, SMPL_V, B, ld_vp, verts, out_scale);
    CKL();
    return 0;
}

extern "C" int smpl_lbs_forward(const float* blend, const float* v_template, const float* j_template, const float* j_shapedirs,
                                const float* skin_weights, const int32_t* parents, const float* pose, const float* betas,
                                const float* trans, int B, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                                void* stream) {
    return smpl_lbs_forward_scaled(blend, nullptr, nullptr, v_template, j_template, j_shapedirs, skin_weights, parents, pose, betas, trans, B,
                                   1.0f, verts, joints, workspace, workspace_bytes, stream);
}

// ---- (f)2: SPIN / HMR ResNet-50 feature extractor (lib/models/spin.py:129-143) -------------------------------------------
namespace {
struct SpinWs {
    SplitOut col, xs, c1s, c2s, sub;      // im2col rows; split copy of the block input; conv1 / conv2 outputs; strided rows
    float *xa, *xb, *rs;                  // block input / output (fp32 NHWC), down-sampled residual
    size_t bytes;
};
SpinWs spin_carve(int B, void* base) {
    SpinWs w;
    Carver c{(char*)base, 0};
    const size_t b = (size_t)B;
    w.col = c.split(b * 12544 * SPIN_STEM_K);      // >= 3136 * 576 (layer1 conv2), 784 * 1152, ...
    w.xs = c.split(b * 3136 * 256);
    w.c1s = c.split(b * 3136 * 128);               // layer2.0 conv1 runs at 56 x 56 with 128 planes
    w.c2s = c.split(b * 3136 * 64);                // = 784 * 256 = 196 * 1024 ...: rows x planes never exceeds 3136 x 64
    w.sub = c.split(b * 784 * 256);
    w.xa = c.f32(b * 3136 * 256); w.xb = c.f32(b * 3136 * 256); w.rs = c.f32(b * 3136 * 256);
    w.bytes = c.cur;
    return w;
}
const int SPIN_PLANES[4] = {64, 128, 256, 512}, SPIN_BLOCKS[4] = {3, 4, 6, 3}, SPIN_STRIDE[4] = {1, 2, 2, 2};
}  // namespace

extern "C" int pmce_spin_num_convs(void) { return 1 + 3 * 16 + 4; }
extern "C" size_t pmce_spin_workspace_bytes(int B) { return B < 1 ? 0 : spin_carve(B, nullptr).bytes; }

extern "C" int pmce_spin_features(const float* w_f32, const void* w_hi, const void* w_lo, const pmce_spin_conv_t* convs, int nconv,
                                  const float* frames, int B, float* feat, void* workspace, size_t workspace_bytes, void* stream) {
    if (!w_f32 || !w_hi || !w_lo || !convs || !frames || !feat || !workspace) { pmce_set_error("pmce_spin_features: NULL argument"); return 2; }
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    if (nconv != pmce_spin_num_convs()) { pmce_set_error("pmce_spin_features: expected %d convolutions, got %d", pmce_spin_num_convs(), nconv); return 2; }
    if (((uintptr_t)workspace & 255) || workspace_bytes < pmce_spin_workspace_bytes(B)) { pmce_set_error("pmce_spin_features: workspace too small or unaligned"); return 2; }
    NvtxRange nvtx("pmce/spin feature extractor");
    cudaStream_t st = (cudaStream_t)stream;
    const SpinWs ws = spin_carve(B, workspace);
    const Weights W{w_f32, (const bf16*)w_hi, (const bf16*)w_lo};
    int ci = 0;
    auto conv = [&](const SplitOut& A, int M, int K, int N, const EpiOpt& o) -> int {
        const pmce_spin_conv_t& c = convs[ci++];
        if (c.cout != N || c.k != K) { pmce_set_error("pmce_spin_features: convolution %d is [%d,%d], expected [%d,%d]", ci - 1, c.cout, c.k, N, K); return 2; }
        EpiOpt e = o;
        e.bias = W.f + c.b_off;
        return linear_tc(A, K, M, K, W, (size_t)c.w_off, K, N, e, st);
    };
    // stem: conv 7x7 / 2 (+ folded BN) -> ReLU -> max-pool 3x3 / 2   (spin.py:131-134)
    {
        const size_t n = (size_t)B * 12544 * (SPIN_STEM_K / 8);
        spin_stem_im2col_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(frames, B, ws.col);
        CKL();
        EpiOpt o; o.out = ws.xb; o.ld_out = 64;
        RET(conv(ws.col, B * 12544, SPIN_STEM_K, 64, o));
        spin_relu_maxpool_kernel<<<cdiv((long long)B * 3136 * 16, 256), 256, 0, st>>>(ws.xb, B, ws.xa, ws.xs);
        CKL();
    }
    float *x = ws.xa, *y = ws.xb;
    int H = 56, Cin = 64;
    for (int li = 0; li < 4; ++li) {
        const int planes = SPIN_PLANES[li];
        for (int bi = 0; bi < SPIN_BLOCKS[li]; ++bi) {
            const int s = bi == 0 ? SPIN_STRIDE[li] : 1;
            const int Ho = H / s, M = B * H * H, Mo = B * Ho * Ho;
            // conv1 1x1 + BN + ReLU   (Bottleneck.forward, spin.py:41-43)
            { EpiOpt o; o.act = 2; o.outs = ws.c1s; o.ld_split = planes; RET(conv(ws.xs, M, Cin, planes, o)); }
            // conv2 3x3 (stride s) + BN + ReLU   (:45-47)
            {
                const size_t n = (size_t)Mo * 9 * (planes / 8);
                spin_im2col3_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(ws.c1s, B, H, H, planes, s, Ho, Ho, ws.col);
                CKL();
                EpiOpt o; o.act = 2; o.outs = ws.c2s; o.ld_split = planes;
                RET(conv(ws.col, Mo, 9 * planes, planes, o));
            }
            // conv3 1x1 + BN, + residual (identity, or the first block's 1x1 / stride s conv + BN, :116-120), ReLU   (:49-55)
            const float* resid = x;
            {
                EpiOpt o; o.out = y; o.ld_out = 4 * planes; o.ld_resid = 4 * planes;
                const int i3 = ci++;                          // conv3 comes before the downsample in the pack order
                if (bi == 0) {
                    SplitOut a = ws.xs;
                    if (s != 1) {
                        const size_t n = (size_t)Mo * (Cin / 8);
                        spin_subsample_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(ws.xs, B, H, H, Cin, s, Ho, Ho, ws.sub);
                        CKL();
                        a = ws.sub;
                    }
                    EpiOpt od; od.out = ws.rs; od.ld_out = 4 * planes;
                    RET(conv(a, Mo, Cin, 4 * planes, od));
                    resid = ws.rs;
                }
                const int after = ci;
                ci = i3;
                o.resid = resid;
                RET(conv(ws.c2s, Mo, planes, 4 * planes, o));
                ci = after > ci ? after : ci;
            }
            spin_relu_split_kernel<<<cdiv((long long)Mo * planes, 256), 256, 0, st>>>(y, (size_t)Mo * planes, ws.xs);
            CKL();
            float* t = x; x = y; y = t;
            H = Ho; Cin = 4 * planes;
        }
    }
    spin_avgpool_kernel<<<cdiv(B * 2048, 256), 256, 0, st>>>(x, B, 2048, feat);
    CKL();
    return 0;
}
