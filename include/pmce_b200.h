/*
 * pmce_b200.h — C ABI of libpmce_b200.so: the B200 (sm_100a) implementation of the PMCE per-clip
 * forward hot path.  Plain C: raw DEVICE pointers (fp32, row-major, contiguous, laid out exactly as the
 * reference's torch tensors), int sizes, a caller-owned workspace and a cudaStream_t passed as void*.
 *
 * Contract:
 *   - every function returns 0 on success, non-zero on error; pmce_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing calls exit().
 *   - the library allocates nothing persistent and owns no buffers: weights blob, workspace, inputs and
 *     outputs all belong to the caller (in the product: the PyTorch caching allocator).
 *   - all work is enqueued on `stream`, asynchronous w.r.t. the host, no device synchronisation, no
 *     allocation => CUDA-graph capturable; re-entrant across streams/devices: concurrent calls on
 *     different streams (each with its own workspace) are independent.  The only process-lifetime
 *     state is the layout cache, per-(kernel, device) function attributes, and one internal side
 *     stream + four events per (device, caller stream), created on first use under a mutex.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef PMCE_B200_H
#define PMCE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMCE_ABI_VERSION 1

/* Model hyper-parameters. Mirrors the constructor arguments / cfg keys the reference reads:
 * models.PMCE.get_model(num_joint, embed_dim, depth) lib/models/PMCE.py:23-26; cfg.DATASET.seqlen,
 * cfg.MODEL.joint_dim/vertx_dim lib/core/config.py:48,61-62. */
typedef struct pmce_dims {
    int32_t num_joint;   /* J: 17 (h36m) or 19 (coco+pelvis+neck)                         */
    int32_t embed_dim;   /* C: lifter width (256 default, 512 supported); multiple of 128 */
    int32_t depth;       /* lifter depth (3)                                             */
    int32_t seqlen;      /* T: frames per clip (16; 64 for the long-clip config)          */
    int32_t num_vert_ds; /* 431 down-sampled vertices                                     */
    int32_t num_vert;    /* 6890 SMPL vertices                                            */
    int32_t feat_dim;    /* 2048 image-feature width                                      */
    int32_t gru_hidden;  /* 1024                                                          */
    int32_t coevo_dim;   /* 64 (joint_dim == vertx_dim)                                   */
    int32_t lifter_heads;/* 8                                                             */
} pmce_dims_t;

/* Where one state_dict tensor lives inside the packed weight blob (units: floats). The tensor, viewed
 * as [rows, cols], is copied to blob[offset + r*ld + c]. */
typedef struct pmce_slot {
    uint64_t offset;
    int64_t rows, cols, ld;
} pmce_slot_t;

const char* pmce_last_error(void);
int pmce_abi_version(void);

/* ---- weights: replaces nn.Module.load_state_dict for the hot path (lib/core/base.py:67) ---------- */
/* Size in bytes of the packed fp32 weight blob for `dims`. */
size_t pmce_weights_bytes(const pmce_dims_t* dims);
/* Look up a reference state_dict key (e.g. "pose_lifter.SpatialBlocks.0.attn.qkv.weight").
 * returns 0 and fills *slot; 1 if the tensor is part of the schema but never reaches an output
 * (the joint branch of coevoblock1/2, lib/models/CoevoDecoder.py:235-236) and is not stored;
 * <0 if the name is not in the schema. */
int pmce_weight_slot(const pmce_dims_t* dims, const char* name, pmce_slot_t* slot);
/* Derived tensors computed on device once all slots are filled: the blob is [fp32 | bf16 hi | bf16 lo] and this
 * writes the split-bf16 copies (hi = bf16(w), lo = bf16(w - hi)) the tensor-core GEMMs read. Must be called
 * after the last slot copy and before any forward entry point. */
int pmce_pack_weights(const pmce_dims_t* dims, void* weights, void* stream);

/* Workspace bytes needed by any forward entry point for batch size B. */
size_t pmce_workspace_bytes(const pmce_dims_t* dims, int B);

/* ---- a1: PMCE.forward, lib/models/PMCE.py:15-20 --------------------------------------------------
 * pose2d [B,T,J,2], img_feat [B,T,2048], vj_relation [431] int32 (nearest joint per vertex,
 * lib/models/CoevoDecoder.py:208,232) -> cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3]. */
int pmce_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat,
                 const int32_t* vj_relation, int B, float* cam_mesh, float* cam_pose, float* pose3d,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- (f)3: overlapping windows of one track (lib/_img_utils.py:58-92 `view_as_windows(indexes, (seqlen,), step=stride)`; the
 * demo runs PMCE.forward once per window with stride 1, main/run_demo.py:145).  pose2d_seq [num_frames,J,2], img_feat_seq
 * [num_frames,2048]; window w = frames [w*stride, w*stride + T), nwin = pmce_sliding_windows(...) = (num_frames-T)/stride + 1;
 * outputs as pmce_forward with B = nwin.  Same results as pmce_forward on the materialised windows, but everything that is a
 * function of one frame is computed once per frame instead of once per window: imgfeat_embed, the token embedding and
 * SpatialBlocks[0] (PoseEstimation.py:78-84), and the GRU layer-0 input projection (CoevoDecoder.py:228).
 * workspace: pmce_workspace_bytes(dims, nwin).  1 <= stride <= T. */
int pmce_sliding_windows(const pmce_dims_t* dims, int num_frames, int stride);
int pmce_forward_sliding(const pmce_dims_t* dims, const void* weights, const float* pose2d_seq, const float* img_feat_seq,
                         const int32_t* vj_relation, int num_frames, int stride, float* cam_mesh, float* cam_pose,
                         float* pose3d, void* workspace, size_t workspace_bytes, void* stream);

/* Same, HOST buffers in and out (pinned or pageable): H2D of the inputs, forward, D2H of the three
 * outputs, then a stream synchronise. d_io must hold pmce_io_bytes(dims,B) device bytes. This is the
 * shape of the call lib/core/base.py:218-238 makes (`.cuda()` ... `.cpu()`). */
size_t pmce_io_bytes(const pmce_dims_t* dims, int B);
int pmce_forward_host(const pmce_dims_t* dims, const void* weights, const float* h_pose2d, const float* h_img_feat,
                      const int32_t* d_vj_relation, int B, float* h_cam_mesh, float* h_cam_pose, float* h_pose3d,
                      void* d_io, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2/a3: GraphormerNet.forward, lib/models/PoseEstimation.py:76-115 -> pose3d [B,J,3] ---------- */
int pmce_lifter_forward(const pmce_dims_t* dims, const void* weights, const float* pose2d, const float* img_feat,
                        int B, float* pose3d, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a4: y[T//2] of nn.GRU(2048,1024,bidirectional,2 layers), lib/models/CoevoDecoder.py:216-221,228-229
 * img_feat [B,T,2048] -> g [B,2048]. Layer 1 only runs the steps y[T//2] depends on. */
int pmce_gru_mid(const pmce_dims_t* dims, const void* weights, const float* img_feat, int B, float* g,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- a5: all live AdaLayerNorm gamma/beta projections at once, lib/models/CoevoDecoder.py:16-29
 * g [B,2048] -> gb [B, pmce_adaln_slots(), 2, 64] (gamma then beta per slot). */
int pmce_adaln_slots(void);
int pmce_adaln_gammabeta(const pmce_dims_t* dims, const void* weights, const float* g, int B, float* gb,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- a8 (with a6 CrossAttentionBlock :64-87 and a7 Block :89-105 inside): CoevoBlock.forward,
 * lib/models/CoevoDecoder.py:175-191.  block in {1,2,3}; joints [B,J,3], verts_in [B,431,3], gb from
 * pmce_adaln_gammabeta -> verts_out [B,431,3]; joints_out [B,J,3] is written only when non-NULL (the
 * reference discards it for blocks 1 and 2, :235-236; weights for it exist only for block 3). */
int pmce_coevo_block(const pmce_dims_t* dims, const void* weights, int block, const float* joints,
                     const float* verts_in, const float* gb, int B, float* joints_out, float* verts_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- a6: CrossAttentionBlock.forward, lib/models/CoevoDecoder.py:82-87 (CrossAttention :47-62, AdaLayerNorm :16-29,
 * timm Mlp).  `which` selects the instance of coevoblock<block>: 0 = joint_CA_FFN (:167; queries = J joints, keys/values =
 * 431 vertices, 8 heads), 1 = vertx_CA_FFN (:169; queries = 431 vertices, keys/values = J joints, 2 heads).
 * xq [B,N1,64], xk / xv [B,N2,64], gb from pmce_adaln_gammabeta -> out [B,N1,64] (may alias xq).  The vertex instance
 * runs the fused kernels: one pass over the query stream for AdaLN_q + Wq + attention + Wp + residual (ca_fused.cuh), one for
 * AdaLN_2 + fc1 + GELU + fc2 + residual with the hidden activations on chip (mlp_fused.cuh); the joint instance uses the
 * second one too. */
int pmce_cross_attn_block(const pmce_dims_t* dims, const void* weights, int block, int which, const float* xq,
                          const float* xk, const float* xv, const float* gb, int B, float* out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Kernel-level entry of the fused vertex cross-attention (measurement / tests; the kernels pmce_cross_attn_block(which=1)
 * and pmce_coevo_block launch): with K / V [B,J,64] = the projected AdaLN'd joint rows (CoevoDecoder.py:53-55),
 *   xq [B,431,64] <- xq + proj(MHA(wq(AdaLN_q(xq)), K, V))   in place                     (:56-62, :84)
 *   t_hi / t_lo [B,431,64] bf16 (optional, both or neither; NULL = skip) <- split-bf16 of AdaLN_2(xq) (:86), by a second launch
 * using the weights of coevoblock<block>.vertx_CA_FFN.  fold_ws: pmce_ca_fold_bytes(B) bytes, 256-byte aligned, holds the
 * per-clip folded operands KQ' [B,48,64] | VP' [B,48,64] (bf16 hi, lo) | sb' [B,48] - AdaLN_q's gamma/beta (from gb), the softmax
 * scale * log2e and the projection bias are folded into them; fold != 0 recomputes them from K / V / gb first (a second,
 * per-clip kernel), fold == 0 reuses what a previous call left there (times the streaming kernel alone). */
size_t pmce_ca_fold_bytes(int B);
int pmce_ca_vertex_fused(const pmce_dims_t* dims, const void* weights, int block, float* xq, const float* K,
                         const float* V, const float* gb, int B, void* t_hi, void* t_lo, void* fold_ws, int fold,
                         void* stream);

/* The same kernel in EMBED mode (what pmce_forward / pmce_coevo_block run under PMCE_CA_EMBED=1; off by default because it
 * measured slower, see api.cu::ca_embed_enabled): the vertex query stream of a block does not exist
 * yet when the block starts - it is  xq[b,i,:] = W_e coords[b,i,:] + b_e + pos[i,:] + Q_embed[i,:]  (CoevoDecoder.py:178,182) -
 * so the kernel takes coords [B,431,3], loads tiles of the per-block table E = b_e + pos + Q_embed [431,64] (shared by all
 * clips, L2-resident) where it would load xq, adds the 3-term product per element, and WRITES xq_out [B,431,64]: no separate
 * embedding launch and no read of the stream.  table_ws: 431*64 floats, 256-byte aligned; fold != 0 (re)computes the folded
 * operands and the table, fold == 0 reuses them.  Same result as embedding first and calling pmce_ca_vertex_fused (up to the
 * rounding order of the embedding's sums). */
int pmce_ca_vertex_fused_embed(const pmce_dims_t* dims, const void* weights, int block, const float* coords, float* xq_out,
                               const float* K, const float* V, const float* gb, int B, void* fold_ws, int fold,
                               float* table_ws, void* stream);

/* ---- a7: Block.forward, lib/models/CoevoDecoder.py:102-105 (Attention :119-131).  `which`: 0 = joint_SA_FFN (:166),
 * 1 = vertx_SA_FFN (:168).  x [B,N,64] -> out [B,N,64] (may alias x). */
int pmce_self_attn_block(const pmce_dims_t* dims, const void* weights, int block, int which, const float* x,
                         const float* gb, int B, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a9 tail: upsample_conv + linear_cur residual, lib/models/CoevoDecoder.py:238-244
 * verts3 [B,431,3], g [B,2048] -> cam_mesh [B,6890,3]. */
int pmce_mesh_epilogue(const pmce_dims_t* dims, const void* weights, const float* verts3, const float* g, int B,
                       float* cam_mesh, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a9: Pose2Mesh.forward, lib/models/CoevoDecoder.py:226-246 (joints already in metres) ---------
 * verts0_out (optional, [B,431,3]) receives the bit-exact gather joints[:, vj_relation, :3] (:232). */
int pmce_decoder_forward(const pmce_dims_t* dims, const void* weights, const float* joints, const float* img_feat,
                         const int32_t* vj_relation, int B, float* cam_pose, float* cam_mesh, float* verts0_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11: J_regressor matvec, lib/core/base.py:225 (`torch.matmul(J_regressor[None], pred_mesh)`) --
 * CSR form of the [R,6890] regressor (<= 11 nnz/row in the shipped arrays): out[b,r,:] =
 * scale * sum_i vals[i] * mesh[b, cols[i], :]. */
int pmce_jregress(const int32_t* row_ptr, const int32_t* cols, const float* vals, int R, const float* mesh,
                  int num_vert, int B, float scale, float* out, void* stream);

/* ---- (f)1, the step after the path: evaluation epilogue of the test loop, lib/core/base.py:223-227 with
 * compute_both_err data/PW3D/dataset.py:269-282 (same code in data/Human36M/dataset.py:611-623).
 * pred_pose[b] = J_regressor . (cam_mesh[b]*scale) (CSR regressor as in pmce_jregress); meshes and joint sets are root-aligned
 * on joint 0; clip_err[b] = (mean over the n_eval evaluation joints, mean over the vertices) of the L2 error of clip b;
 * mean_err = their means over the batch = the reference's (joint_mean_error, mesh_mean_error).
 * cam_mesh, gt_mesh [B,V,3] in metres (scaled by `scale`, 1000 in the reference); gt_pose [B,R,3] and pred_pose [B,R,3] in mm;
 * eval_joints [n_eval] int32 rows of the regressor (data/PW3D/dataset.py:35).  Replaces four D2H copies of [B,6890,3]
 * tensors + numpy per batch by a 2-float result. */
int pmce_eval_errors(const int32_t* row_ptr, const int32_t* cols, const float* vals, int R, const float* cam_mesh,
                     const float* gt_mesh, const float* gt_pose, const int32_t* eval_joints, int n_eval, int num_vert,
                     int B, float scale, float* pred_pose, float* clip_err, float* mean_err, void* stream);

/* ---- nn.Linear forward as used by every projection on the path (F.linear; e.g. lib/models/CoevoDecoder.py:19-20,
 * timm Mlp fc1/fc2): out[M,N] = act(x[M,K] weight[N,K]^T + bias[N]); act 0 = none, 1 = exact GELU. K % 4 == 0.
 * Exposed so the dominant GEMM can be timed / profiled in isolation. */
int pmce_linear(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                void* stream);

/* Same contract on the tcgen05 tensor-core path: operands are split on device into bf16 hi/lo pairs
 * (3 bf16 MMAs per product, fp32 accumulate in TMEM, ~2^-16 relative error), tiles fed by TMA. K % 8 == 0.
 * scratch: pmce_linear_tc_scratch_bytes(M,N,K) device bytes, 256-byte aligned. */
size_t pmce_linear_tc_scratch_bytes(int M, int N, int K);
int pmce_linear_tc(const float* x, const float* weight, const float* bias, int M, int N, int K, int act, float* out,
                   void* scratch, size_t scratch_bytes, void* stream);
/* The two halves of pmce_linear_tc, so the GEMM kernel can be timed alone: split fp32 [rows,cols] (cols % 4 == 0)
 * into bf16 hi/lo (uint16 storage), and the GEMM on already split operands (x_hi/x_lo [M,K], w_hi/w_lo [N,K]) with the
 * epilogue variants the forward uses: fp32 out (+ residual), or GELU + split-bf16 out. */
int pmce_split_bf16(const float* x, int rows, int cols, void* hi, void* lo, void* stream);
int pmce_linear_tc_presplit(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                            int M, int N, int K, int act, float* out /* fp32 [M,N] or NULL */,
                            void* out_hi, void* out_lo /* split bf16 [M,N] or NULL */, const float* resid /* [M,N] or NULL */,
                            void* stream);

/* ---- (f)2, the step before the path: the SPIN / HMR ResNet-50 feature extractor, lib/models/spin.py:129-143
 * (`HMR.feature_extractor`, called on the frame crops at main/run_demo.py:315) -> the 2048-d per-frame features PMCE.forward
 * consumes.  frames [B,3,224,224] fp32 NCHW -> feat [B,2048].  Convolutions run as NHWC GEMMs on the tensor cores; BatchNorm
 * (eval) is folded at pack time.  `convs`: HOST array of the 53 convolutions in execution order (stem; per bottleneck conv1,
 * conv2, conv3, then the first block's downsample), each a [cout, k] matrix (k = kh*kw*cin in (kh, kw, cin) order, the stem's
 * 147 padded to 152) at float offset w_off of the three weight views (fp32 / bf16 hi / bf16 lo) with its folded bias at b_off of
 * the fp32 view.  pmce_b200/spin.py builds them from the reference state_dict. */
typedef struct pmce_spin_conv {
    int64_t w_off, b_off;
    int32_t cout, k;
} pmce_spin_conv_t;
int pmce_spin_num_convs(void);
size_t pmce_spin_workspace_bytes(int B);
int pmce_spin_features(const float* w_f32, const void* w_hi, const void* w_lo, const pmce_spin_conv_t* convs, int nconv,
                       const float* frames, int B, float* feat, void* workspace, size_t workspace_bytes, void* stream);

/* Cumulative number of kernels this library has launched in this process (for the bench's gpu_launches). */
unsigned long long pmce_launch_count(void);

/* ---- a12: SMPL_Layer.forward, smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158 ----------------
 * blend  [20670, KB] fp32, KB = smpl_blend_ld() >= 217: row (v*3+c) = [shapedirs[v,c,0:10] | posedirs[v,c,0:207] | 0]
 * v_template [20670]; j_template [24,3] = Jreg @ v_template; j_shapedirs [24,3,10] = Jreg @ shapedirs;
 * skin_weights [6890,24]; parents [24] int32; pose [B,72]; betas [B,10]; trans [B,3] or NULL
 * -> verts [B,6890,3], joints [B,24,3]. workspace: smpl_workspace_bytes(B). */
int smpl_blend_ld(void);
size_t smpl_workspace_bytes(int B);
int smpl_lbs_forward(const float* blend, const float* v_template, const float* j_template, const float* j_shapedirs,
                     const float* skin_weights, const int32_t* parents, const float* pose, const float* betas,
                     const float* trans, int B, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                     void* stream);

/* (f)4, the data path's use of the layer: get_smpl_coord (data/PW3D/dataset.py:70-88, data/Human36M/dataset.py:91-110 ...)
 * runs SMPL_Layer.forward per sample on the CPU inside the DataLoader workers and multiplies by 1000 (metre -> millimetre).
 * Same as smpl_lbs_forward for a whole batch, with verts and joints multiplied by out_scale on the way out.  When
 * blend_hi / blend_lo are given (split-bf16 copies of the blend matrix from pmce_split_bf16, [20672, smpl_blend_ld()] with
 * rows >= 20670 zero, and v_template padded to 20672 floats) the blend-shape product runs on the tensor cores (bf16x3) and
 * `blend` may be NULL; otherwise it is the exact fp32 CUDA-core GEMM. */
int smpl_lbs_forward_scaled(const float* blend, const void* blend_hi, const void* blend_lo, const float* v_template,
                            const float* j_template, const float* j_shapedirs, const float* skin_weights,
                            const int32_t* parents, const float* pose, const float* betas, const float* trans, int B,
                            float out_scale, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Same with the skinning weights in sparse form: skin_idx4 / skin_w4 [6890,4] = the (joint, weight) pairs of each vertex in
 * ascending joint order, zero-weight padded (every shipped SMPL model binds a vertex to <= 4 joints; smpl_layer.py:134 sums
 * over all 24, the other 20 terms being exact zeros). NULL for both selects the dense [6890,24] skin_weights. */
int smpl_lbs_forward_sparse(const float* blend, const void* blend_hi, const void* blend_lo, const float* v_template,
                            const float* j_template, const float* j_shapedirs, const float* skin_weights, const int32_t* skin_idx4,
                            const float* skin_w4, const int32_t* parents, const float* pose, const float* betas, const float* trans,
                            int B, float out_scale, float* verts, float* joints, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMCE_B200_H */
