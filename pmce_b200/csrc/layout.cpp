// Packed-weight layout construction and state_dict-name lookup (host only).
#include "layout.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>
#include <memory>

static thread_local char g_err[512] = "";

void pmce_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pmce_last_error(void) { return g_err; }
extern "C" int pmce_abi_version(void) { return PMCE_ABI_VERSION; }
extern "C" int pmce_adaln_slots(void) { return PMCE_ADALN_SLOTS; }

namespace {

struct Builder {
    Layout* L;
    size_t cur = 0;
    size_t alloc(size_t n) {
        size_t off = cur;
        cur += (n + 63) / 64 * 64;   // 256-byte aligned slots
        return off;
    }
    size_t add(const std::string& name, int64_t rows, int64_t cols, int64_t ld = -1) {
        if (ld < 0) ld = cols;
        size_t off = alloc((size_t)rows * ld);
        put(name, off, rows, cols, ld);
        return off;
    }
    void put(const std::string& name, size_t off, int64_t rows, int64_t cols, int64_t ld) {
        pmce_slot_t s;
        s.offset = off; s.rows = rows; s.cols = cols; s.ld = ld;
        L->slots[name] = s;
    }
    void dead(const std::string& name) { L->dead[name] = 1; }
    // nn.Linear: weight [out,in], bias [out]
    void lin(const std::string& p, int64_t out, int64_t in, size_t* w, size_t* b) {
        *w = add(p + ".weight", out, in);
        *b = add(p + ".bias", 1, out);
    }
    void dead_lin(const std::string& p) { dead(p + ".weight"); dead(p + ".bias"); }
};

std::string fmt(const char* f, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return std::string(buf);
}

Layout* build_layout(const pmce_dims_t& d) {
    Layout* L = new Layout();
    L->d = d;
    Builder b;
    b.L = L;
    const int64_t J = d.num_joint, C = d.embed_dim, T = d.seqlen, Vd = d.num_vert_ds, V = d.num_vert, F = d.feat_dim,
                  H = d.gru_hidden, D = d.coevo_dim;

    // ---- lifter (lib/models/PoseEstimation.py:31-66) ----
    std::string p = "pose_lifter.";
    b.lin(p + "joint_embed", C, 2, &L->jew, &L->jeb);
    b.lin(p + "imgfeat_embed", C, F, &L->iew, &L->ieb);
    L->spos = b.add(p + "spatial_pos_embed", J, C);
    L->tpos = b.add(p + "temporal_pos_embed", T, C);
    for (int kind = 0; kind < 2; ++kind)
        for (int i = 0; i < d.depth; ++i) {
            VitBlockW& w = kind == 0 ? L->sp[i] : L->tp[i];
            std::string q = p + fmt("%s.%d.", kind == 0 ? "SpatialBlocks" : "TemporalBlocks", i);
            w.n1w = b.add(q + "norm1.weight", 1, C); w.n1b = b.add(q + "norm1.bias", 1, C);
            b.lin(q + "attn.qkv", 3 * C, C, &w.qkvw, &w.qkvb);
            b.lin(q + "attn.proj", C, C, &w.projw, &w.projb);
            w.n2w = b.add(q + "norm2.weight", 1, C); w.n2b = b.add(q + "norm2.bias", 1, C);
            b.lin(q + "mlp.fc1", 2 * C, C, &w.fc1w, &w.fc1b);
            b.lin(q + "mlp.fc2", C, 2 * C, &w.fc2w, &w.fc2b);
        }
    L->nsw = b.add(p + "norm_s.weight", 1, C); L->nsb = b.add(p + "norm_s.bias", 1, C);
    L->ntw = b.add(p + "norm_t.weight", 1, C); L->ntb = b.add(p + "norm_t.bias", 1, C);
    L->r0w = b.add(p + "regression.0.weight", 1, C); L->r0b = b.add(p + "regression.0.bias", 1, C);
    b.lin(p + "regression.1", 3, C, &L->r1w, &L->r1b);
    L->fusw = b.add(p + "fusion.weight", 1, T); L->fusb = b.add(p + "fusion.bias", 1, 1);

    // ---- decoder (lib/models/CoevoDecoder.py:193-224) ----
    p = "pose_mesh_coevo.";
    L->init_vertices = b.add(p + "init_vertices", Vd, 3);
    L->adaln_w = b.alloc((size_t)PMCE_ADALN_SLOTS * 2 * D * F);
    L->adaln_b = b.alloc((size_t)PMCE_ADALN_SLOTS * 2 * D);
    int next_slot = 0;
    auto adaln = [&](const std::string& q, bool alive) -> int {
        if (!alive) {
            b.dead_lin(q + ".mlp_gamma"); b.dead_lin(q + ".mlp_beta");
            return -1;
        }
        const int s = next_slot++;
        b.put(q + ".mlp_gamma.weight", L->adaln_w + (size_t)(s * 2 * D) * F, D, F, F);
        b.put(q + ".mlp_gamma.bias", L->adaln_b + (size_t)(s * 2 * D), 1, D, D);
        b.put(q + ".mlp_beta.weight", L->adaln_w + (size_t)(s * 2 * D + D) * F, D, F, F);
        b.put(q + ".mlp_beta.bias", L->adaln_b + (size_t)(s * 2 * D + D), 1, D, D);
        return s;
    };
    auto opt_lin = [&](const std::string& q, int64_t out, int64_t in, size_t* w, size_t* bb, bool alive) {
        if (alive) b.lin(q, out, in, w, bb);
        else { b.dead_lin(q); *w = *bb = (size_t)-1; }
    };
    auto opt_add = [&](const std::string& q, int64_t rows, int64_t cols, bool alive) -> size_t {
        if (alive) return b.add(q, rows, cols);
        b.dead(q);
        return (size_t)-1;
    };
    for (int k = 0; k < 3; ++k) {
        CoevoW& w = L->blk[k];
        // joints1/joints2 are discarded by Pose2Mesh.forward (:235-236): the joint branch only matters in block 3
        const bool ja = (k == 2);
        w.joint_alive = ja;
        std::string q = p + fmt("coevoblock%d.", k + 1);
        w.jpos = b.add(q + "joint_pos_embed", J, D);
        w.jQ = opt_add(q + "j_Q_embed", J, D, ja);
        w.j2vK = b.add(q + "j2v_K_embed", J, D);
        w.vpos = b.add(q + "vertx_pos_embed", Vd, D);
        w.vQ = b.add(q + "v_Q_embed", Vd, D);
        w.v2jK = opt_add(q + "v2j_K_embed", Vd, D, ja);
        b.lin(q + "joint_proj", D, 3, &w.jprojw, &w.jprojb);
        b.lin(q + "vertx_proj", D, 3, &w.vprojw, &w.vprojb);
        opt_lin(q + "proj_v2j_dim", D, D, &w.v2jw, &w.v2jb, ja);
        b.lin(q + "proj_j2v_dim", D, D, &w.j2vw, &w.j2vb);
        for (int st = 0; st < 2; ++st) {   // 0 = vertx first (slot order), 1 = joint
            const bool alive = st == 0 ? true : ja;
            const char* nm = st == 0 ? "vertx" : "joint";
            CaW& ca = st == 0 ? w.vca : w.jca;
            SaW& sa = st == 0 ? w.vsa : w.jsa;
            std::string c = q + fmt("%s_CA_FFN.", nm);
            ca.sq = adaln(c + "normq", alive); ca.sk = adaln(c + "normk", alive); ca.sv = adaln(c + "normv", alive);
            opt_lin(c + "attn.wq", D, D, &ca.wq, &ca.bq, alive);
            opt_lin(c + "attn.wk", D, D, &ca.wk, &ca.bk, alive);
            opt_lin(c + "attn.wv", D, D, &ca.wv, &ca.bv, alive);
            opt_lin(c + "attn.proj", D, D, &ca.wp, &ca.bp, alive);
            ca.s2 = adaln(c + "norm2", alive);
            opt_lin(c + "mlp.fc1", 4 * D, D, &ca.fc1w, &ca.fc1b, alive);
            opt_lin(c + "mlp.fc2", D, 4 * D, &ca.fc2w, &ca.fc2b, alive);
            std::string s = q + fmt("%s_SA_FFN.", nm);
            sa.s1 = adaln(s + "norm1", alive);
            opt_lin(s + "attn.qkv", 3 * D, D, &sa.qkvw, &sa.qkvb, alive);
            opt_lin(s + "attn.proj", D, D, &sa.wp, &sa.bp, alive);
            sa.s2 = adaln(s + "norm2", alive);
            opt_lin(s + "mlp.fc1", 4 * D, D, &sa.fc1w, &sa.fc1b, alive);
            opt_lin(s + "mlp.fc2", D, 4 * D, &sa.fc2w, &sa.fc2b, alive);
        }
        opt_lin(q + "proj_joint_feat2coor", 3, D, &w.jf2cw, &w.jf2cb, ja);
        b.lin(q + "proj_vertx_feat2coor", 3, D, &w.vf2cw, &w.vf2cb);
    }
    if (next_slot != PMCE_ADALN_SLOTS) { delete L; pmce_set_error("internal: adaln slot count %d", next_slot); return nullptr; }

    L->ups_ld = (int)((Vd * 3 + 7) / 8 * 8);   // K padded for 16-byte TMA row strides
    L->ups_w = b.add(p + "upsample_conv.weight", V, Vd * 3, L->ups_ld);
    L->ups_b = b.add(p + "upsample_conv.bias", 1, V);

    // GRU (nn.GRU(2048,1024,bidirectional,num_layers=2), :216-221)
    L->wih0 = b.alloc((size_t)6 * H * F);
    L->bih0 = b.alloc((size_t)6 * H);
    for (int dir = 0; dir < 2; ++dir) {
        const char* sfx = dir == 0 ? "" : "_reverse";
        b.put(p + fmt("gru_cur.weight_ih_l0%s", sfx), L->wih0 + (size_t)dir * 3 * H * F, 3 * H, F, F);
        b.put(p + fmt("gru_cur.bias_ih_l0%s", sfx), L->bih0 + (size_t)dir * 3 * H, 1, 3 * H, 3 * H);
        L->whh0[dir] = b.add(p + fmt("gru_cur.weight_hh_l0%s", sfx), 3 * H, H);
        L->bhh0[dir] = b.add(p + fmt("gru_cur.bias_hh_l0%s", sfx), 1, 3 * H);
        L->wih1[dir] = b.add(p + fmt("gru_cur.weight_ih_l1%s", sfx), 3 * H, 2 * H);
        L->bih1[dir] = b.add(p + fmt("gru_cur.bias_ih_l1%s", sfx), 1, 3 * H);
        L->whh1[dir] = b.add(p + fmt("gru_cur.weight_hh_l1%s", sfx), 3 * H, H);
        L->bhh1[dir] = b.add(p + fmt("gru_cur.bias_hh_l1%s", sfx), 1, 3 * H);
    }
    L->lc_w = b.alloc((size_t)3 * V * 2 * H);
    L->lc_b = b.alloc((size_t)3 * V);
    for (int i = 0; i < 3; ++i) {
        b.put(p + fmt("linear_cur%d.weight", i + 1), L->lc_w + (size_t)i * V * 2 * H, V, 2 * H, 2 * H);
        b.put(p + fmt("linear_cur%d.bias", i + 1), L->lc_b + (size_t)i * V, 1, V, V);
    }
    L->total_floats = b.cur;
    return L;
}

std::mutex g_mu;
std::vector<std::unique_ptr<Layout>> g_layouts;

bool dims_ok(const pmce_dims_t& d) {
    if (d.num_joint < 1 || d.num_joint > 64) { pmce_set_error("num_joint %d out of range [1,64]", d.num_joint); return false; }
    if (d.embed_dim < 128 || d.embed_dim > 1024 || d.embed_dim % 128) { pmce_set_error("embed_dim %d must be 128, 256 or 512", d.embed_dim); return false; }
    if (d.lifter_heads != 8 || (d.embed_dim / d.lifter_heads != 16 && d.embed_dim / d.lifter_heads != 32 && d.embed_dim / d.lifter_heads != 64)) { pmce_set_error("unsupported lifter head_dim %d", d.embed_dim / (d.lifter_heads ? d.lifter_heads : 1)); return false; }
    if (d.depth < 1 || d.depth > PMCE_MAX_DEPTH) { pmce_set_error("depth %d out of range", d.depth); return false; }
    if (d.seqlen < 1 || d.seqlen > 256) { pmce_set_error("seqlen %d out of range [1,256]", d.seqlen); return false; }
    if (d.coevo_dim != 64) { pmce_set_error("coevo_dim must be 64 (got %d)", d.coevo_dim); return false; }
    if (d.gru_hidden % 64 || d.gru_hidden < 64) { pmce_set_error("gru_hidden must be a multiple of 64"); return false; }
    if (d.feat_dim != 2 * d.gru_hidden) { pmce_set_error("feat_dim must equal 2*gru_hidden (AdaLN/linear_cur consume y[T//2])"); return false; }
    if (d.feat_dim % 4) { pmce_set_error("feat_dim must be a multiple of 4"); return false; }
    if (d.num_vert_ds < 1 || d.num_vert < 1) { pmce_set_error("bad vertex counts"); return false; }
    return true;
}

}  // namespace

const Layout* pmce_get_layout(const pmce_dims_t* dims) {
    if (!dims) { pmce_set_error("dims is NULL"); return nullptr; }
    if (!dims_ok(*dims)) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& l : g_layouts)
        if (memcmp(&l->d, dims, sizeof(pmce_dims_t)) == 0) return l.get();
    Layout* L = build_layout(*dims);
    if (!L) return nullptr;
    g_layouts.emplace_back(L);
    return L;
}

extern "C" int pmce_weight_slot(const pmce_dims_t* dims, const char* name, pmce_slot_t* slot) {
    const Layout* L = pmce_get_layout(dims);
    if (!L) return -2;
    if (!name || !slot) { pmce_set_error("NULL argument"); return -2; }
    auto it = L->slots.find(name);
    if (it != L->slots.end()) { *slot = it->second; return 0; }
    if (L->dead.count(name)) { memset(slot, 0, sizeof(*slot)); return 1; }
    pmce_set_error("unknown state_dict key '%s'", name);
    return -1;
}
