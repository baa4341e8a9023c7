// Packed-weight layout: maps the reference state_dict schema (431 tensors) onto one fp32 blob.
// Packing is placement only: tensors that the kernels consume as one matrix are placed adjacently
//   * all live AdaLayerNorm gamma/beta projections  -> adaln_w [slots*128, 2048], adaln_b [slots*128]
//   * linear_cur1..3                                  -> lc_w [3*6890, 2048], lc_b [3*6890]
//   * GRU layer-0 input weights fwd|bwd               -> wih0 [6144, 2048]
//   * upsample_conv.weight [6890, 431*3] with the row stride padded to a multiple of 4 floats
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <map>
#include <string>
#include "../../include/pmce_b200.h"

struct VitBlockW { size_t n1w, n1b, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b; };
struct CaW { size_t wq, bq, wk, bk, wv, bv, wp, bp, fc1w, fc1b, fc2w, fc2b; int sq, sk, sv, s2; };
struct SaW { size_t qkvw, qkvb, wp, bp, fc1w, fc1b, fc2w, fc2b; int s1, s2; };
struct CoevoW {
    size_t jpos, jQ, j2vK, vpos, vQ, v2jK;
    size_t jprojw, jprojb, vprojw, vprojb, v2jw, v2jb, j2vw, j2vb;
    SaW jsa, vsa;
    CaW jca, vca;
    size_t jf2cw, jf2cb, vf2cw, vf2cb;
    bool joint_alive;
};

#define PMCE_MAX_DEPTH 8
#define PMCE_ADALN_SLOTS 24

struct Layout {
    pmce_dims_t d;
    size_t total_floats;
    std::map<std::string, pmce_slot_t> slots;   // live tensors
    std::map<std::string, int> dead;            // schema members that never reach an output
    // lifter
    size_t jew, jeb, iew, ieb, spos, tpos, nsw, nsb, ntw, ntb, r0w, r0b, r1w, r1b, fusw, fusb;
    VitBlockW sp[PMCE_MAX_DEPTH], tp[PMCE_MAX_DEPTH];
    // decoder
    size_t init_vertices;
    CoevoW blk[3];
    size_t adaln_w, adaln_b;
    size_t ups_w, ups_b; int ups_ld;
    size_t wih0, bih0;              // [6144,2048], [6144]  (fwd rows then bwd rows)
    size_t whh0[2], bhh0[2];        // [3072,1024]
    size_t wih1[2], bih1[2], whh1[2], bhh1[2];
    size_t lc_w, lc_b;
};

const Layout* pmce_get_layout(const pmce_dims_t* dims);   // cached; nullptr + error message on invalid dims
void pmce_set_error(const char* fmt, ...);
