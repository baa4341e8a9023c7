// libpmce_b200: C-ABI entry points and host-side orchestration of the PMCE forward hot path.
// All device work is enqueued on the caller's stream; no allocation, no synchronisation (graph-capturable).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "../../include/pmce_b200.h"
#include "layout.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "gemm_tc.cuh"

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
#define RET(x) do { int _r = (x); if (_r) return _r; } while (0)

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long pmce_launch_count(void) { return g_launches.load(); }

namespace {

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------------
// workspace carving (floats, 256-byte aligned chunks)
// ---------------------------------------------------------------------------------------------------
struct Carver {
    float* base;
    size_t cur;
    float* take(size_t n) {
        float* p = base ? base + cur : nullptr;
        cur += (n + 63) / 64 * 64;
        return p;
    }
};

struct Workspace {
    // lifter
    float *imgemb, *x, *xn, *qkv, *att, *hid, *r3, *joints_m;
    // gru
    float *gi0, *y0, *gi1f, *gi1b, *h1[2][2], *g;
    // decoder
    float *gb, *verts[3], *Jf, *Vf, *xqv, *tA, *tA2, *Qv, *xkj, *tJ, *Kj, *Vj, *att_d, *qkv_d, *hid_d;
    float *xqj, *xkv, *Kv, *Vv, *Qj, *attj, *qkvj, *hidj, *im2col;
    size_t floats;
};

Workspace carve(const pmce_dims_t& d, int B, void* base) {
    Workspace w;
    Carver c{(float*)base, 0};
    const size_t J = d.num_joint, C = d.embed_dim, T = d.seqlen, Vd = d.num_vert_ds, H = d.gru_hidden, F = d.feat_dim, D = d.coevo_dim;
    const size_t N = (size_t)B * T * J;
    w.imgemb = c.take((size_t)B * T * C);
    w.x = c.take(N * C); w.xn = c.take(N * C); w.qkv = c.take(N * 3 * C); w.att = c.take(N * C); w.hid = c.take(N * 2 * C);
    w.r3 = c.take(N * 3); w.joints_m = c.take((size_t)B * J * 3);
    const size_t mid = T / 2;
    w.gi0 = c.take((size_t)T * B * 6 * H); w.y0 = c.take((size_t)T * B * 2 * H);
    w.gi1f = c.take((mid + 1) * B * 3 * H); w.gi1b = c.take((T - mid) * B * 3 * H);
    for (int dir = 0; dir < 2; ++dir) for (int i = 0; i < 2; ++i) w.h1[dir][i] = c.take((size_t)B * H);
    w.g = c.take((size_t)B * F);
    w.gb = c.take((size_t)B * PMCE_ADALN_SLOTS * 2 * D);
    for (int i = 0; i < 3; ++i) w.verts[i] = c.take((size_t)B * Vd * 3);
    const size_t nv = (size_t)B * Vd, nj = (size_t)B * J;
    w.Jf = c.take(nj * D); w.Vf = c.take(nv * D); w.xqv = c.take(nv * D); w.tA = c.take(nv * D); w.tA2 = c.take(nv * D);
    w.Qv = c.take(nv * D); w.xkj = c.take(nj * D); w.tJ = c.take(nj * D); w.Kj = c.take(nj * D); w.Vj = c.take(nj * D);
    w.att_d = c.take(nv * D); w.qkv_d = c.take(nv * 3 * D); w.hid_d = c.take(nv * 4 * D);
    w.xqj = c.take(nj * D); w.xkv = c.take(nv * D); w.Kv = c.take(nv * D); w.Vv = c.take(nv * D); w.Qj = c.take(nj * D);
    w.attj = c.take(nj * D); w.qkvj = c.take(nj * 3 * D); w.hidj = c.take(nj * 4 * D);
    w.im2col = c.take((size_t)B * 3 * ((Vd * 3 + 3) / 4 * 4));
    w.floats = c.cur;
    return w;
}

int check_ws(const pmce_dims_t& d, int B, void* ws, size_t bytes, Workspace* out) {
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    if (!ws) { pmce_set_error("workspace is NULL"); return 2; }
    if (((uintptr_t)ws) & 255) { pmce_set_error("workspace must be 256-byte aligned"); return 2; }
    *out = carve(d, B, ws);
    if (out->floats * sizeof(float) > bytes) {
        pmce_set_error("workspace too small: need %zu bytes, got %zu", out->floats * sizeof(float), bytes);
        return 2;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
int linear(const float* A, int lda, const float* W, int ldw, const float* bias, float* out, int ldc, int M, int N, int K,
           cudaStream_t st, int act = 0, const float* resid = nullptr, const float* rowadd = nullptr, int period = 1) {
    GemmEpi e = gemm_epi_plain(ldc);
    e.bias = bias; e.act = act; e.resid = resid; e.rowadd = rowadd; e.rowadd_period = period;
    CKG(launch_gemm_tn(A, lda, W, ldw, out, M, N, K, e, st));
    return 0;
}

template <int D>
int launch_attn_d(const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, float* O, AttnAddr ao, int nseq,
                  int H, int N1, int N2, cudaStream_t st) {
    const size_t smem = (size_t)N2 * D * 2 * sizeof(float);
    static size_t configured = 0;   // per-process, per-instantiation
    if (smem > 48 * 1024 && smem > configured) {
        CK(cudaFuncSetAttribute(attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = 227 * 1024;
    }
    if (smem > 227 * 1024) { pmce_set_error("attention K/V tile (%zu B) exceeds shared memory", smem); return 3; }
    const int threads = N1 >= 128 ? 128 : (N1 > 64 ? 96 : (N1 > 32 ? 64 : 32));
    dim3 grid(cdiv(N1, threads), H, nseq);
    if (nseq > 65535) { pmce_set_error("too many attention sequences (%d) for one launch", nseq); return 3; }
    attn_kernel<D><<<grid, threads, smem, st>>>(Q, aq, K, V, akv, O, ao, N1, N2, 1.0f / sqrtf((float)D));
    CKL();
    return 0;
}

int launch_attn(int D, const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, float* O, AttnAddr ao,
                int nseq, int H, int N1, int N2, cudaStream_t st) {
    // split very large sequence counts into several launches (gridDim.z limit)
    switch (D) {
        case 8: return launch_attn_d<8>(Q, aq, K, V, akv, O, ao, nseq, H, N1, N2, st);
        case 16: return launch_attn_d<16>(Q, aq, K, V, akv, O, ao, nseq, H, N1, N2, st);
        case 32: return launch_attn_d<32>(Q, aq, K, V, akv, O, ao, nseq, H, N1, N2, st);
        case 64: return launch_attn_d<64>(Q, aq, K, V, akv, O, ao, nseq, H, N1, N2, st);
    }
    pmce_set_error("unsupported head_dim %d", D);
    return 3;
}

AttnAddr addr_plain(int ntok, int ld) {
    AttnAddr a;
    a.seq.div = 1; a.seq.s0 = ntok; a.seq.s1 = 0; a.tok = 1; a.ld = ld;
    return a;
}

int ln_rows(const float* x, int nrows, int C, const LnParams* a, const float* pos, int pos_div, int pos_mod, float* out1,
            const LnParams* b, float* out2, cudaStream_t st) {
    LnParams za{nullptr, nullptr, 0.f};
    ln_rows_kernel<<<cdiv(nrows, 8), 256, 0, st>>>(x, nrows, C, a ? *a : za, a ? 1 : 0, pos, pos_div, pos_mod, out1,
                                                     b ? *b : za, b ? out2 : nullptr);
    CKL();
    return 0;
}

int adaln(const float* x, int B, int ntok, const float* gb, int slot, float* y, cudaStream_t st) {
    const int nrows = B * ntok;
    adaln_apply_kernel<<<cdiv(nrows, 8), 256, 0, st>>>(x, nrows, ntok, gb, PMCE_ADALN_SLOTS * 128, slot, 1e-6f, y);
    CKL();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a2/a3 lifter
// ---------------------------------------------------------------------------------------------------
int vit_block(const pmce_dims_t& d, const float* Wt, const VitBlockW& w, const Workspace& ws, int B, bool temporal, cudaStream_t st) {
    const int J = d.num_joint, C = d.embed_dim, T = d.seqlen, Hh = d.lifter_heads;
    const int N = B * T * J;
    // qkv
    RET(linear(ws.xn, C, Wt + w.qkvw, C, Wt + w.qkvb, ws.qkv, 3 * C, N, 3 * C, C, st));
    AttnAddr a, ao;
    int nseq, L;
    if (!temporal) { a.seq.div = 1; a.seq.s0 = J; a.seq.s1 = 0; a.tok = 1; nseq = B * T; L = J; }
    else { a.seq.div = J; a.seq.s0 = (long long)T * J; a.seq.s1 = 1; a.tok = J; nseq = B * J; L = T; }
    a.ld = 3 * C;
    ao = a; ao.ld = C;
    // gridDim.z <= 65535: chunk over sequences in multiples of J (keeps the (b,j) decomposition intact)
    const int chunk = (65535 / J) * J;
    for (int s0 = 0; s0 < nseq; s0 += chunk) {
        const int ns = nseq - s0 < chunk ? nseq - s0 : chunk;
        const long long r0 = a.seq(s0);   // first row of sequence s0 (s0 % J == 0 in temporal mode)
        RET(launch_attn(C / Hh, ws.qkv + r0 * 3 * C, a, ws.qkv + r0 * 3 * C + C, ws.qkv + r0 * 3 * C + 2 * C, a,
                        ws.att + r0 * C, ao, ns, Hh, L, L, st));
    }
    // proj + residual (in place on x)
    RET(linear(ws.att, C, Wt + w.projw, C, Wt + w.projb, ws.x, C, N, C, C, st, 0, ws.x));
    // norm2 -> xn
    LnParams n2{Wt + w.n2w, Wt + w.n2b, 1e-6f};
    RET(ln_rows(ws.x, N, C, &n2, nullptr, 1, 1, ws.xn, nullptr, nullptr, st));
    RET(linear(ws.xn, C, Wt + w.fc1w, C, Wt + w.fc1b, ws.hid, 2 * C, N, 2 * C, C, st, 1));
    RET(linear(ws.hid, 2 * C, Wt + w.fc2w, 2 * C, Wt + w.fc2b, ws.x, C, N, C, 2 * C, st, 0, ws.x));
    return 0;
}

int lifter(const Layout& L, const float* Wt, const float* pose2d, const float* img_feat, int B, float* pose3d,
           const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int J = d.num_joint, C = d.embed_dim, T = d.seqlen, F = d.feat_dim;
    const int N = B * T * J;
    RET(linear(img_feat, F, Wt + L.iew, F, Wt + L.ieb, ws.imgemb, C, B * T, C, F, st));
    LnParams s0n1{Wt + L.sp[0].n1w, Wt + L.sp[0].n1b, 1e-6f};
    lifter_embed_kernel<<<cdiv(N, 8), 256, 0, st>>>(pose2d, ws.imgemb, Wt + L.jew, Wt + L.jeb, Wt + L.spos, N, J, C, s0n1, ws.x, ws.xn);
    CKL();
    LnParams ns{Wt + L.nsw, Wt + L.nsb, 1e-6f}, nt{Wt + L.ntw, Wt + L.ntb, 1e-6f};
    for (int i = 0; i < d.depth; ++i) {
        RET(vit_block(d, Wt, L.sp[i], ws, B, false, st));
        LnParams tn1{Wt + L.tp[i].n1w, Wt + L.tp[i].n1b, 1e-6f};
        RET(ln_rows(ws.x, N, C, &ns, i == 0 ? Wt + L.tpos : nullptr, J, T, ws.x, &tn1, ws.xn, st));
        RET(vit_block(d, Wt, L.tp[i], ws, B, true, st));
        if (i + 1 < d.depth) {
            LnParams sn1{Wt + L.sp[i + 1].n1w, Wt + L.sp[i + 1].n1b, 1e-6f};
            RET(ln_rows(ws.x, N, C, &nt, nullptr, 1, 1, ws.x, &sn1, ws.xn, st));
        }
    }
    LnParams nh{Wt + L.r0w, Wt + L.r0b, 1e-5f};
    lifter_head_kernel<<<cdiv(N, 8), 256, 0, st>>>(ws.x, N, C, nt, nh, Wt + L.r1w, Wt + L.r1b, ws.r3);
    CKL();
    lifter_fuse_kernel<<<cdiv(B * J * 3, 256), 256, 0, st>>>(ws.r3, Wt + L.fusw, Wt + L.fusb, B, T, J, pose3d, ws.joints_m);
    CKL();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a4 GRU
// ---------------------------------------------------------------------------------------------------
int gru_step(const GruDir* dirs, int ndir, int B, int H, cudaStream_t st) {
    dim3 grid(H / 16, cdiv(B, 64), ndir);
    gru_step_kernel<<<grid, 256, 0, st>>>(dirs[0], dirs[ndir > 1 ? 1 : 0], B, H);
    CKL();
    return 0;
}

int gru_mid(const Layout& L, const float* Wt, const float* img_feat, int B, float* g, const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int T = d.seqlen, H = d.gru_hidden, F = d.feat_dim;
    const int mid = T / 2;
    // layer-0 input projections for every frame and both directions, written time-major: gi0[t][b][6H]
    {
        GemmEpi e = gemm_epi_plain(6 * H);
        e.bias = Wt + L.bih0;
        e.rmap.div = T; e.rmap.s0 = 6 * H; e.rmap.s1 = (long long)B * 6 * H;   // row (b,t) -> t*B*6H + b*6H
        CKG(launch_gemm_tn(img_feat, F, Wt + L.wih0, F, ws.gi0, B * T, 6 * H, F, e, st));
    }
    for (int s = 0; s < T; ++s) {
        GruDir dd[2];
        const int tf = s, tb = T - 1 - s;
        dd[0].gi = ws.gi0 + (size_t)tf * B * 6 * H; dd[0].ld_gi = 6 * H;
        dd[0].hprev = s > 0 ? ws.y0 + (size_t)(tf - 1) * B * 2 * H : nullptr; dd[0].ld_h = 2 * H;
        dd[0].whh = Wt + L.whh0[0]; dd[0].bhh = Wt + L.bhh0[0];
        dd[0].hout = ws.y0 + (size_t)tf * B * 2 * H; dd[0].ld_o = 2 * H; dd[0].hout2 = nullptr; dd[0].ld_o2 = 0;
        dd[1].gi = ws.gi0 + (size_t)tb * B * 6 * H + 3 * H; dd[1].ld_gi = 6 * H;
        dd[1].hprev = s > 0 ? ws.y0 + (size_t)(tb + 1) * B * 2 * H + H : nullptr; dd[1].ld_h = 2 * H;
        dd[1].whh = Wt + L.whh0[1]; dd[1].bhh = Wt + L.bhh0[1];
        dd[1].hout = ws.y0 + (size_t)tb * B * 2 * H + H; dd[1].ld_o = 2 * H; dd[1].hout2 = nullptr; dd[1].ld_o2 = 0;
        RET(gru_step(dd, 2, B, H, st));
    }
    // layer 1: only the steps y[T//2] depends on (fwd t = 0..mid, bwd t = T-1..mid)
    const int nf = mid + 1, nb = T - mid;
    RET(linear(ws.y0, 2 * H, Wt + L.wih1[0], 2 * H, Wt + L.bih1[0], ws.gi1f, 3 * H, nf * B, 3 * H, 2 * H, st));
    RET(linear(ws.y0 + (size_t)mid * B * 2 * H, 2 * H, Wt + L.wih1[1], 2 * H, Wt + L.bih1[1], ws.gi1b, 3 * H, nb * B, 3 * H, 2 * H, st));
    const int nsteps = nf > nb ? nf : nb;
    for (int s = 0; s < nsteps; ++s) {
        GruDir dd[2];
        int n = 0;
        if (s < nf) {
            GruDir& x = dd[n++];
            x.gi = ws.gi1f + (size_t)s * B * 3 * H; x.ld_gi = 3 * H;
            x.hprev = s > 0 ? ws.h1[0][(s - 1) & 1] : nullptr; x.ld_h = H;
            x.whh = Wt + L.whh1[0]; x.bhh = Wt + L.bhh1[0];
            if (s == nf - 1) { x.hout = g; x.ld_o = 2 * H; } else { x.hout = ws.h1[0][s & 1]; x.ld_o = H; }
            x.hout2 = nullptr; x.ld_o2 = 0;
        }
        if (s < nb) {
            GruDir& x = dd[n++];
            const int t = T - 1 - s;
            x.gi = ws.gi1b + (size_t)(t - mid) * B * 3 * H; x.ld_gi = 3 * H;
            x.hprev = s > 0 ? ws.h1[1][(s - 1) & 1] : nullptr; x.ld_h = H;
            x.whh = Wt + L.whh1[1]; x.bhh = Wt + L.bhh1[1];
            if (s == nb - 1) { x.hout = g + H; x.ld_o = 2 * H; } else { x.hout = ws.h1[1][s & 1]; x.ld_o = H; }
            x.hout2 = nullptr; x.ld_o2 = 0;
        }
        RET(gru_step(dd, n, B, H, st));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a5 AdaLN gamma/beta, a6-a8 co-evolution block
// ---------------------------------------------------------------------------------------------------
int adaln_gammabeta(const Layout& L, const float* Wt, const float* g, int B, float* gb, cudaStream_t st) {
    const int N = PMCE_ADALN_SLOTS * 2 * L.d.coevo_dim, F = L.d.feat_dim;
    return linear(g, F, Wt + L.adaln_w, F, Wt + L.adaln_b, gb, N, B, N, F, st);
}

// residual-stream tail shared by CrossAttentionBlock and Block: x += proj(att); x += fc2(gelu(fc1(AdaLN_2(x))))
int attn_tail(const float* Wt, size_t wp, size_t bp, int s2, size_t fc1w, size_t fc1b, size_t fc2w, size_t fc2b, float* x,
              const float* att, float* tmp, float* hid, const float* gb, int B, int ntok, cudaStream_t st) {
    const int M = B * ntok;
    RET(linear(att, 64, Wt + wp, 64, Wt + bp, x, 64, M, 64, 64, st, 0, x));
    RET(adaln(x, B, ntok, gb, s2, tmp, st));
    RET(linear(tmp, 64, Wt + fc1w, 64, Wt + fc1b, hid, 256, M, 256, 64, st, 1));
    RET(linear(hid, 256, Wt + fc2w, 256, Wt + fc2b, x, 64, M, 64, 256, st, 0, x));
    return 0;
}

int coevo_block(const Layout& L, const float* Wt, int k, const float* joints, const float* verts_in, const float* gb, int B,
                float* joints_out, float* verts_out, const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const CoevoW& w = L.blk[k];
    const int J = d.num_joint, Vd = d.num_vert_ds;
    const int nj = B * J, nv = B * Vd;
    const bool ja = joints_out != nullptr;
    if (ja && !w.joint_alive) { pmce_set_error("coevoblock%d: joint-branch weights are not stored (output is discarded by the reference)", k + 1); return 4; }

    // coordinate -> feature (+pos), query streams (+Q embed)   (CoevoDecoder.py:177-183)
    coevo_embed_kernel<<<cdiv((long long)nj * 16, 256), 256, 0, st>>>(joints, nj, J, Wt + w.jprojw, Wt + w.jprojb, Wt + w.jpos,
                                                                      ja ? Wt + w.jQ : nullptr, ws.Jf, ja ? ws.xqj : nullptr);
    CKL();
    coevo_embed_kernel<<<cdiv((long long)nv * 16, 256), 256, 0, st>>>(verts_in, nv, Vd, Wt + w.vprojw, Wt + w.vprojb, Wt + w.vpos,
                                                                      Wt + w.vQ, ja ? ws.Vf : nullptr, ws.xqv);
    CKL();
    // keys: proj_j2v(Jf) + j2v_K ; proj_v2j(Vf) + v2j_K  — both from the PRE-update features (:183-184)
    RET(linear(ws.Jf, 64, Wt + w.j2vw, 64, Wt + w.j2vb, ws.xkj, 64, nj, 64, 64, st, 0, nullptr, Wt + w.j2vK, J));
    if (ja) RET(linear(ws.Vf, 64, Wt + w.v2jw, 64, Wt + w.v2jb, ws.xkv, 64, nv, 64, 64, st, 0, nullptr, Wt + w.v2jK, Vd));

    // ---- vertex cross-attention block: q = vertices (431), k/v = joints (J); 2 heads x 32 ----
    RET(adaln(ws.xqv, B, Vd, gb, w.vca.sq, ws.tA, st));
    RET(linear(ws.tA, 64, Wt + w.vca.wq, 64, Wt + w.vca.bq, ws.Qv, 64, nv, 64, 64, st));
    RET(adaln(ws.xkj, B, J, gb, w.vca.sk, ws.tJ, st));
    RET(linear(ws.tJ, 64, Wt + w.vca.wk, 64, Wt + w.vca.bk, ws.Kj, 64, nj, 64, 64, st));
    RET(adaln(ws.Jf, B, J, gb, w.vca.sv, ws.tJ, st));
    RET(linear(ws.tJ, 64, Wt + w.vca.wv, 64, Wt + w.vca.bv, ws.Vj, 64, nj, 64, 64, st));
    RET(launch_attn(32, ws.Qv, addr_plain(Vd, 64), ws.Kj, ws.Vj, addr_plain(J, 64), ws.att_d, addr_plain(Vd, 64), B, 2, Vd, J, st));
    RET(attn_tail(Wt, w.vca.wp, w.vca.bp, w.vca.s2, w.vca.fc1w, w.vca.fc1b, w.vca.fc2w, w.vca.fc2b, ws.xqv, ws.att_d, ws.tA,
                  ws.hid_d, gb, B, Vd, st));

    if (ja) {
        // ---- joint cross-attention block: q = joints (J), k/v = vertices (431); 8 heads x 8 ----
        RET(adaln(ws.xqj, B, J, gb, w.jca.sq, ws.tJ, st));
        RET(linear(ws.tJ, 64, Wt + w.jca.wq, 64, Wt + w.jca.bq, ws.Qj, 64, nj, 64, 64, st));
        RET(adaln(ws.xkv, B, Vd, gb, w.jca.sk, ws.tA, st));
        RET(linear(ws.tA, 64, Wt + w.jca.wk, 64, Wt + w.jca.bk, ws.Kv, 64, nv, 64, 64, st));
        RET(adaln(ws.Vf, B, Vd, gb, w.jca.sv, ws.tA2, st));
        RET(linear(ws.tA2, 64, Wt + w.jca.wv, 64, Wt + w.jca.bv, ws.Vv, 64, nv, 64, 64, st));
        RET(launch_attn(8, ws.Qj, addr_plain(J, 64), ws.Kv, ws.Vv, addr_plain(Vd, 64), ws.attj, addr_plain(J, 64), B, 8, J, Vd, st));
        RET(attn_tail(Wt, w.jca.wp, w.jca.bp, w.jca.s2, w.jca.fc1w, w.jca.fc1b, w.jca.fc2w, w.jca.fc2b, ws.xqj, ws.attj, ws.tJ,
                      ws.hidj, gb, B, J, st));
        // ---- joint self-attention block ----
        RET(adaln(ws.xqj, B, J, gb, w.jsa.s1, ws.tJ, st));
        RET(linear(ws.tJ, 64, Wt + w.jsa.qkvw, 64, Wt + w.jsa.qkvb, ws.qkvj, 192, nj, 192, 64, st));
        RET(launch_attn(8, ws.qkvj, addr_plain(J, 192), ws.qkvj + 64, ws.qkvj + 128, addr_plain(J, 192), ws.attj, addr_plain(J, 64), B, 8, J, J, st));
        RET(attn_tail(Wt, w.jsa.wp, w.jsa.bp, w.jsa.s2, w.jsa.fc1w, w.jsa.fc1b, w.jsa.fc2w, w.jsa.fc2b, ws.xqj, ws.attj, ws.tJ,
                      ws.hidj, gb, B, J, st));
        feat2coor_kernel<<<cdiv(nj, 8), 256, 0, st>>>(ws.xqj, nj, Wt + w.jf2cw, Wt + w.jf2cb, joints, joints_out);
        CKL();
    }

    // ---- vertex self-attention block: 431 x 431, 2 heads x 32 ----
    RET(adaln(ws.xqv, B, Vd, gb, w.vsa.s1, ws.tA, st));
    RET(linear(ws.tA, 64, Wt + w.vsa.qkvw, 64, Wt + w.vsa.qkvb, ws.qkv_d, 192, nv, 192, 64, st));
    RET(launch_attn(32, ws.qkv_d, addr_plain(Vd, 192), ws.qkv_d + 64, ws.qkv_d + 128, addr_plain(Vd, 192), ws.att_d, addr_plain(Vd, 64), B, 2, Vd, Vd, st));
    RET(attn_tail(Wt, w.vsa.wp, w.vsa.bp, w.vsa.s2, w.vsa.fc1w, w.vsa.fc1b, w.vsa.fc2w, w.vsa.fc2b, ws.xqv, ws.att_d, ws.tA,
                  ws.hid_d, gb, B, Vd, st));
    feat2coor_kernel<<<cdiv(nv, 8), 256, 0, st>>>(ws.xqv, nv, Wt + w.vf2cw, Wt + w.vf2cb, verts_in, verts_out);
    CKL();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// a9 tail: upsample_conv + linear_cur residual
// ---------------------------------------------------------------------------------------------------
int mesh_epilogue(const Layout& L, const float* Wt, const float* verts3, const float* g, int B, float* mesh, const Workspace& ws,
                  cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int Vd = d.num_vert_ds, V = d.num_vert, F = d.feat_dim, ldk = L.ups_ld;
    upsample_im2col_kernel<<<cdiv((long long)B * 3 * ldk, 256), 256, 0, st>>>(verts3, B, Vd, ldk, ws.im2col);
    CKL();
    {   // mesh[b,o,l] = b_up[o] + sum_{c,k} W_up[o,c,k] verts3[b,c,l+k-1]
        GemmEpi e = gemm_epi_plain(0);
        e.bias = Wt + L.ups_b;
        e.rmap.div = 3; e.rmap.s0 = (long long)V * 3; e.rmap.s1 = 1;
        e.cmap.div = 1; e.cmap.s0 = 3; e.cmap.s1 = 0;
        CKG(launch_gemm_tn(ws.im2col, ldk, Wt + L.ups_w, ldk, mesh, B * 3, V, ldk, e, st));
    }
    {   // mesh[b,o,l] += W_cur{l+1} relu(g[b]) + b_cur{l+1}
        GemmEpi e = gemm_epi_plain(0);
        e.bias = Wt + L.lc_b; e.a_relu = 1; e.resid = mesh;
        e.rmap.div = 1; e.rmap.s0 = (long long)V * 3; e.rmap.s1 = 0;
        e.cmap.div = V; e.cmap.s0 = 1; e.cmap.s1 = 3;
        CKG(launch_gemm_tn(g, F, Wt + L.lc_w, F, mesh, B, 3 * V, F, e, st));
    }
    return 0;
}

int decoder(const Layout& L, const float* Wt, const float* joints, const float* img_feat, const int32_t* vj, int B, float* cam_pose,
            float* cam_mesh, float* verts0_out, const Workspace& ws, cudaStream_t st) {
    const pmce_dims_t& d = L.d;
    const int J = d.num_joint, Vd = d.num_vert_ds;
    RET(gru_mid(L, Wt, img_feat, B, ws.g, ws, st));
    RET(adaln_gammabeta(L, Wt, ws.g, B, ws.gb, st));
    float* v0 = verts0_out ? verts0_out : ws.verts[2];
    gather_verts_kernel<<<cdiv((long long)B * Vd * 3, 256), 256, 0, st>>>(joints, vj, B, J, Vd, v0);
    CKL();
    RET(coevo_block(L, Wt, 0, joints, v0, ws.gb, B, nullptr, ws.verts[0], ws, st));
    RET(coevo_block(L, Wt, 1, joints, ws.verts[0], ws.gb, B, nullptr, ws.verts[1], ws, st));
    RET(coevo_block(L, Wt, 2, joints, ws.verts[1], ws.gb, B, cam_pose, ws.verts[0], ws, st));
    RET(mesh_epilogue(L, Wt, ws.verts[0], ws.g, B, cam_mesh, ws, st));
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
    const float* Wt = (const float*)weights;           \
    if (!Wt) { pmce_set_error("weights is NULL"); return 2; } \
    cudaStream_t st = (cudaStream_t)stream;

extern "C" size_t pmce_workspace_bytes(const pmce_dims_t* dims, int B) {
    const Layout* Lp = pmce_get_layout(dims);
    if (!Lp || B < 1) return 0;
    return carve(*dims, B, nullptr).floats * sizeof(float);
}

extern "C" int pmce_pack_weights(const pmce_dims_t* dims, void* weights, void* stream) {
    (void)stream;
    const Layout* Lp = pmce_get_layout(dims);
    if (!Lp) return 1;
    if (!weights) { pmce_set_error("weights is NULL"); return 2; }
    return 0;
}

extern "C" int pmce_lifter_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat, int B,
                                   float* pose3d, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return lifter(L, Wt, pose2d, img_feat, B, pose3d, ws, st);
}

extern "C" int pmce_gru_mid(const pmce_dims_t* dims, const void* weights, const float* img_feat, int B, float* g, void* workspace,
                            size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return gru_mid(L, Wt, img_feat, B, g, ws, st);
}

extern "C" int pmce_adaln_gammabeta(const pmce_dims_t* dims, const void* weights, const float* g, int B, float* gb, void* stream) {
    GET_LAYOUT();
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    return adaln_gammabeta(L, Wt, g, B, gb, st);
}

extern "C" int pmce_coevo_block(const pmce_dims_t* dims, const void* weights, int block, const float* joints, const float* verts_in,
                                const float* gb, int B, float* joints_out, float* verts_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
    GET_LAYOUT();
    if (block < 1 || block > 3) { pmce_set_error("block must be 1..3 (got %d)", block); return 2; }
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return coevo_block(L, Wt, block - 1, joints, verts_in, gb, B, joints_out, verts_out, ws, st);
}

extern "C" int pmce_mesh_epilogue(const pmce_dims_t* dims, const void* weights, const float* verts3, const float* g, int B,
                                  float* cam_mesh, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return mesh_epilogue(L, Wt, verts3, g, B, cam_mesh, ws, st);
}

extern "C" int pmce_decoder_forward(const pmce_dims_t* dims, const void* weights, const float* joints, const float* img_feat,
                                    const int32_t* vj_relation, int B, float* cam_pose, float* cam_mesh, float* verts0_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    return decoder(L, Wt, joints, img_feat, vj_relation, B, cam_pose, cam_mesh, verts0_out, ws, st);
}

extern "C" int pmce_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat,
                            const int32_t* vj_relation, int B, float* cam_mesh, float* cam_pose, float* pose3d, void* workspace,
                            size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    Workspace ws;
    RET(check_ws(*dims, B, workspace, workspace_bytes, &ws));
    RET(lifter(L, Wt, pose2d, img_feat, B, pose3d, ws, st));
    return decoder(L, Wt, ws.joints_m, img_feat, vj_relation, B, cam_pose, cam_mesh, nullptr, ws, st);
}

extern "C" size_t pmce_io_bytes(const pmce_dims_t* dims, int B) {
    if (!dims || B < 1) return 0;
    const size_t J = dims->num_joint, T = dims->seqlen, F = dims->feat_dim, V = dims->num_vert;
    auto al = [](size_t n) { return (n + 63) / 64 * 64; };
    return (al((size_t)B * T * J * 2) + al((size_t)B * T * F) + al((size_t)B * V * 3) + 2 * al((size_t)B * J * 3)) * sizeof(float);
}

extern "C" int pmce_forward_host(const pmce_dims_t* dims, const void* weights, const float* h_pose2d, const float* h_img_feat,
                                 const int32_t* d_vj_relation, int B, float* h_cam_mesh, float* h_cam_pose, float* h_pose3d,
                                 void* d_io, void* workspace, size_t workspace_bytes, void* stream) {
    GET_LAYOUT();
    (void)L;
    if (!d_io) { pmce_set_error("d_io is NULL"); return 2; }
    const size_t J = dims->num_joint, T = dims->seqlen, F = dims->feat_dim, V = dims->num_vert;
    Carver c{(float*)d_io, 0};
    float* d_p2d = c.take((size_t)B * T * J * 2);
    float* d_feat = c.take((size_t)B * T * F);
    float* d_mesh = c.take((size_t)B * V * 3);
    float* d_pose = c.take((size_t)B * J * 3);
    float* d_p3d = c.take((size_t)B * J * 3);
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

extern "C" int pmce_linear(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                           void* stream) {
    if (!x || !weight || !out || M < 1 || N < 1 || K < 4 || (K & 3)) { pmce_set_error("pmce_linear: bad argument (K must be a multiple of 4)"); return 2; }
    return linear(x, K, weight, K, bias, out, N, M, N, K, (cudaStream_t)stream, act);
}

extern "C" size_t pmce_linear_tc_scratch_bytes(int M, int N, int K) {
    auto al = [](size_t n) { return (n + 127) / 128 * 128; };
    return 2 * (al((size_t)M * K) + al((size_t)N * K)) * sizeof(__nv_bfloat16);
}

extern "C" int pmce_linear_tc(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                              void* scratch, size_t scratch_bytes, void* stream) {
    if (!x || !weight || !out || !scratch || M < 1 || N < 1 || K < 8 || (K & 7)) { pmce_set_error("pmce_linear_tc: bad argument (K must be a multiple of 8)"); return 2; }
    if (scratch_bytes < pmce_linear_tc_scratch_bytes(M, N, K)) { pmce_set_error("pmce_linear_tc: scratch too small"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    auto al = [](size_t n) { return (n + 127) / 128 * 128; };
    __nv_bfloat16* a_hi = (__nv_bfloat16*)scratch;
    __nv_bfloat16* a_lo = a_hi + al((size_t)M * K);
    __nv_bfloat16* w_hi = a_lo + al((size_t)M * K);
    __nv_bfloat16* w_lo = w_hi + al((size_t)N * K);
    split_rows_kernel<<<cdiv((long long)M * (K / 4), 256), 256, 0, st>>>(x, M, K, K, 0, a_hi, a_lo, K);
    CKL();
    split_rows_kernel<<<cdiv((long long)N * (K / 4), 256), 256, 0, st>>>(weight, N, K, K, 0, w_hi, w_lo, K);
    CKL();
    TcOperand A{a_hi, a_lo, M, K, K}, W{w_hi, w_lo, N, K, K};
    TcEpi e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.act = act; e.out_f32 = out; e.ld_out = N; e.rowadd_period = 1;
    int rc = launch_linear_tc(A, W, e, st);
    count_launch();
    if (rc) { pmce_set_error("pmce_linear_tc: launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return 10; }
    return 0;
}

// ---- SMPL LBS --------------------------------------------------------------------------------------
#define SMPL_V 6890
#define SMPL_LDK 220
extern "C" int smpl_blend_ld(void) { return SMPL_LDK; }
extern "C" size_t smpl_workspace_bytes(int B) {
    if (B < 1) return 0;
    auto al = [](size_t n) { return (n + 63) / 64 * 64; };
    return (al((size_t)B * SMPL_LDK) + al((size_t)B * 288) + al((size_t)B * SMPL_V * 3)) * sizeof(float);
}

extern "C" int smpl_lbs_forward(const float* blend, const float* v_template, const float* j_template, const float* j_shapedirs,
                                const float* skin_weights, const int32_t* parents, const float* pose, const float* betas,
                                const float* trans, int B, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!blend || !v_template || !j_template || !j_shapedirs || !skin_weights || !parents || !pose || !betas || !verts || !joints) {
        pmce_set_error("smpl_lbs_forward: NULL argument"); return 2;
    }
    if (B < 1) { pmce_set_error("batch size %d < 1", B); return 2; }
    if (!workspace || workspace_bytes < smpl_workspace_bytes(B)) { pmce_set_error("smpl workspace too small"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    Carver c{(float*)workspace, 0};
    float* coef = c.take((size_t)B * SMPL_LDK);
    float* Amat = c.take((size_t)B * 288);
    float* vposed = c.take((size_t)B * SMPL_V * 3);
    smpl_pose_kernel<<<B, 32, 0, st>>>(pose, betas, trans, j_template, j_shapedirs, parents, B, SMPL_LDK, coef, Amat, joints);
    CKL();
    // v_posed[b, v*3+c] = v_template + [shapedirs | posedirs] . [betas | pose_map]   (smpl_layer.py:93-99)
    RET(linear(coef, SMPL_LDK, blend, SMPL_LDK, v_template, vposed, SMPL_V * 3, B, SMPL_V * 3, SMPL_LDK, st));
    dim3 grid(cdiv(SMPL_V, 256), B);
    smpl_skin_kernel<<<grid, 256, 0, st>>>(vposed, Amat, skin_weights, trans, SMPL_V, verts);
    CKL();
    return 0;
}
