/*
 * slotvps_b200 -- C ABI of the B200-native (sm_100a) Slot-VPS retriever hot path.
 *
 * The reference has no FFI for this path: it is PyTorch module code.  Each entry point below
 * replaces one reference interface (cited as file:line under /root/reference) and is what a
 * ctypes / torch-extension binding on the reference side would call (INTEGRATION.md shows the
 * stub).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to fp32 unless noted;
 *   - every function takes the CUDA stream it must enqueue on (cudaStream_t passed as void*),
 *     never allocates, never synchronises the device unless stated; scratch memory comes from
 *     the caller (slotvps_*_workspace_bytes);
 *   - returns 0 on success, a negative SLOTVPS_E* code otherwise; slotvps_last_error() gives
 *     a thread-local message.  Thread-compatible, not re-entrant on one workspace.
 *   - tensors use the reference's own layouts: features NCHW ([C][h][w] per frame), slots
 *     [N][C] row-major, id maps [H][W] int64.
 */
#ifndef SLOTVPS_B200_H_
#define SLOTVPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLOTVPS_MAX_LEVELS 4
#define SLOTVPS_MAX_STAGES 8
#define SLOTVPS_MAX_FRAMES 8
#define SLOTVPS_C 256              /* dh_dim, fixed by the kernels */
#define SLOTVPS_CIN 128            /* channels of the per-level input features */

enum {
  SLOTVPS_OK = 0,
  SLOTVPS_EINVAL = -1,             /* bad argument / unsupported shape */
  SLOTVPS_ECUDA = -2,              /* a CUDA call or launch failed */
  SLOTVPS_EWORKSPACE = -3,         /* workspace too small */
  SLOTVPS_EUNSUPPORTED = -4        /* device is not sm_100 */
};

/* ---- per-stage parameter table --------------------------------------------------------
 * One entry per MaskRCNNHead stage (dynamic_mask_head.py:231-289), pointers into the module's
 * own parameter storage (state_dict keys head_series_{l}.{j}.<name>); tq_* are the
 * temporal_query_head.* parameters (TemporalSlotsHead, :465-492) or NULL when the stage has
 * no Video Retriever.  Weights are row-major [out][in] as nn.Linear stores them.            */
typedef struct slotvps_stage_params {
  const float *in_proj_w, *in_proj_b;          /* self_attn.in_proj_{weight,bias} [768,256],[768] */
  const float *out_proj_w, *out_proj_b;        /* self_attn.out_proj.*            [256,256],[256] */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b, *norm3_w, *norm3_b;
  const float *to_q_w, *to_q_b, *to_k_w, *to_k_b, *to_v_w, *to_v_b;   /* inst_interact.to_* */
  const float *nq_w, *nq_b, *nk_w, *nk_b, *nv_w, *nv_b, *no_w, *no_b; /* inst_interact.norm_{q,k,v}, norm1 */
  const float *lin1_w, *lin1_b, *lin2_w, *lin2_b;                     /* [F,256],[F] / [256,F],[256] */
  const float *cls0_w, *cls0_nw, *cls0_nb, *cls1_w, *cls1_nw, *cls1_nb; /* cls_module.{0,1,3,4} */
  const float *reg0_w, *reg0_nw, *reg0_nb, *reg1_w, *reg1_nw, *reg1_nb; /* reg_module.{0,1,3,4} */
  const float *logit_w, *logit_b;                                      /* class_logits [K,256],[K] */
  /* Video Retriever (NULL if absent) */
  const float *tq_to_q_w, *tq_to_q_b, *tq_to_k_w, *tq_to_k_b, *tq_to_v_w, *tq_to_v_b;
  const float *tq_nq_w, *tq_nq_b, *tq_nk_w, *tq_nk_b, *tq_nv_w, *tq_nv_b, *tq_no_w, *tq_no_b;
  const float *tq_lin1_w, *tq_lin1_b, *tq_lin2_w, *tq_lin2_b;
  const float *tq_norm2_w, *tq_norm2_b, *tq_norm3_w, *tq_norm3_b;
} slotvps_stage_params;

/* Shape of one head invocation == the kwargs of MultiScaleDynamicMaskHead.__init__
 * (dynamic_mask_head.py:38-54) that change the computation, plus the input sizes. */
typedef struct slotvps_head_desc {
  int32_t n_frames;                            /* T = len(features)                         */
  int32_t n_slots;                             /* N = proposal_num (<= 512)                 */
  int32_t n_levels;                            /* feat_num_levels (<= 4)                    */
  int32_t heads_per_level[SLOTVPS_MAX_LEVELS]; /* per_dh_num_heads                          */
  int32_t h[SLOTVPS_MAX_LEVELS], w[SLOTVPS_MAX_LEVELS]; /* coarse -> fine, each 2x the previous */
  int32_t num_classes;                         /* 20                                        */
  int32_t dim_feedforward;                     /* 2048                                      */
  int32_t temporal_dim_feedforward;            /* 1024                                      */
  int32_t nhead;                               /* 8                                         */
  int32_t temporal_mask;                       /* bit s set: stage s runs the Video Retriever */
  int32_t pos_mode;                            /* 0: no pos, 1: pos tensors given, 2: sine embedding generated on the fly */
  int32_t kernel_path;                         /* 0: auto (tcgen05 when shapes allow), 1: force fp32 CUDA-core path */
  int32_t ffn_act;                             /* stage FFN activation (`activation`, r50_fpn_slotvps.py:33 / swinL_fpn_slotvps.py:41):
                                                  0 = the r50 config's "gelu" (exact erf form), 1 = "relu", 2 = "gelu"          */
  int32_t temporal_ffn_act;                    /* Video Retriever FFN activation (temporal_query_attention_config.activation,
                                                  r50 :49 "relu" / swinL :56 "gelu"): 0 = "relu" (r50), 1 = "relu", 2 = "gelu"   */
} slotvps_head_desc;

/* Bytes of scratch slotvps_head_forward needs for this shape. */
int slotvps_head_workspace_bytes(const slotvps_head_desc* d, size_t* bytes);

/* One-time (per weight set) preparation: folds/centres/splits the stage weights into the
 * operand formats of the kernels.  `prepared` is a caller-owned device buffer of
 * slotvps_prepared_bytes() bytes that must outlive the forward calls using it.             */
int slotvps_prepared_bytes(const slotvps_head_desc* d, size_t* bytes);
int slotvps_prepare_weights(const slotvps_head_desc* d, const slotvps_stage_params* stages /*[n_stages]*/,
                            const float* conv_trans_w /*[256,384]*/, const float* conv_trans_b /*[256]*/,
                            void* prepared, void* stream);
/* Same, with the 1x1 input transform of the caller folded into the level fusion:
 * semantic_trans_ins applies VPS_Capsule.conv_trans (Conv2d(128,128,1) + bias, no norm / activation,
 * vps_capsule.py:74-79, vps_temporal_slots.py:129-135) to every level before the head.  With
 * in_trans_w [128,128] / in_trans_b [128] given, slotvps_head_forward takes the UN-transformed semantic-head
 * features and (W_b T) f + (W_b t + b) replaces W_b (T f + t) + b, which removes that pass (5.7 GFLOP and one
 * read + write of every input level per frame).  NULL / NULL = slotvps_prepare_weights.                  */
int slotvps_prepare_weights_ex(const slotvps_head_desc* d, const slotvps_stage_params* stages, const float* conv_w,
                               const float* conv_b, const float* in_trans_w, const float* in_trans_b, void* prepared,
                               void* stream);

/* MultiScaleDynamicMaskHead.forward (dynamic_mask_head.py:138-228), bs == 1.
 *   feats[t*n_levels+l] : [128,h_l,w_l]     input features (the reference's features[t][l][0])
 *   pos[t*n_levels+l]   : [256,h_l,w_l]     position embedding, or NULL array when pos_mode != 1
 *   init_query[t]       : [N,256]
 *   cls_out             : [T][S][N][num_classes]   (the reference's T x [S,1,N,num_classes])
 *   emb_out             : [T][S][N][256]
 *   fused_out[t*n_levels+l] : [256,h_l,w_l]  transformed features (third return value)      */
int slotvps_head_forward(const slotvps_head_desc* d, const slotvps_stage_params* stages,
                         const void* prepared,
                         const float* const* feats, const float* const* pos,
                         const float* const* init_query,
                         float* cls_out, float* emb_out, float* const* fused_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* slotvps_head_forward with options.  Every field may be 0 / NULL (= slotvps_head_forward).
 *   stage_slots_in[s*T+t] : [N,256] or NULL.  Teacher forcing for per-stage parity: the slots ENTERING stage s of frame t
 *                           are taken from here instead of from stage s-1 (dynamic_mask_head.py:210-211 carries them).
 *   skip_fused_mask       : bit l set = the fp32 NCHW feature of level l (third return value of the reference head) is
 *                           not written; only allowed for levels that run the tensor-core path (their fp16 operand
 *                           planes feed everything downstream, incl. slotvps_head_mask_logits).  The L2 integration /
 *                           SlotVPSRetriever set it: simple_test only consumes the finest level, through the planes.
 *   feat_bn_scale/shift   : [256] folded feat_bn (eval) of generate_final_outputs (vps_temporal_slots.py:145-149).  When
 *                           given, the finest level's fusion epilogue also accumulates sum_c (scale*x+shift)^2 per pixel
 *                           into rnorm_ss [4][T][P_last] (four 64-channel partial sums), which slotvps_head_mask_logits_ex then uses
 *                           instead of re-reading the fp32 feature.                                                     */
typedef struct slotvps_head_opts {
  const float* const* stage_slots_in;
  int32_t skip_fused_mask;
  const float* feat_bn_scale;
  const float* feat_bn_shift;
  float* rnorm_ss;
} slotvps_head_opts;
int slotvps_head_forward_ex(const slotvps_head_desc* d, const slotvps_stage_params* stages,
                            const void* prepared,
                            const float* const* feats, const float* const* pos,
                            const float* const* init_query,
                            float* cls_out, float* emb_out, float* const* fused_out,
                            void* workspace, size_t workspace_bytes, const slotvps_head_opts* opts, void* stream);

/* Level fusion alone (dynamic_mask_head.py:172-185): out = conv_trans(cat(up2x(prev), x)) or,
 * prev == NULL, conv_trans(cat(x,x,x)).  prev [256,h/2,w/2], x [128,h,w], out [256,h,w];
 * scratch: 256*max(128,(h/2)*(w/2)) floats.                                                 */
int slotvps_level_fuse(const float* prev, const float* x, const float* conv_w, const float* conv_b,
                       float* out, int h, int w, float* scratch, void* stream);

/* Panoptic Retriever attention alone (MaskDynamicConv.forward, dynamic_mask_head.py:423-461),
 * one frame: slots_p [N,256] (post-norm1 slots), x [256,h,w], pos [256,h,w] | NULL
 * -> out [N,256] = ReLU(LN(sum_pixels softmax_slots(q k^T) v)).                              */
int slotvps_slot_attention(const slotvps_stage_params* stage, const float* slots_p, const float* x,
                           const float* pos, float* out, int n_slots, int h, int w,
                           int kernel_path, void* workspace, size_t workspace_bytes, void* stream);
int slotvps_slot_attention_workspace_bytes(int n_slots, int h, int w, size_t* bytes);

/* Mask-logit projection (VPS_Temporal_Slots.generate_final_outputs, vps_temporal_slots.py:144-154):
 *   feat [256,h,w] finest fused feature, emb [N,256], feat_bn_* [256] (BatchNorm2d eval),
 *   fg_bn = {weight, bias, running_mean, running_var} of BatchNorm2d(1) -> out [N,h,w].      */
int slotvps_mask_logits_workspace_bytes(int n_slots, int h, int w, size_t* bytes);
int slotvps_mask_logits(const float* feat, const float* emb,
                        const float* feat_bn_w, const float* feat_bn_b,
                        const float* feat_bn_mean, const float* feat_bn_var,
                        const float* fg_bn /*[4] device: weight,bias,running_mean,running_var*/, float* out,
                        int n_slots, int h, int w, void* workspace, size_t workspace_bytes, void* stream);

/* Same projection, but the contraction reads the fp16 operand planes that the immediately preceding
 * slotvps_head_forward(d, ..., head_workspace) left for the finest level (frame = index into the clip);
 * `feat` (the same [256,h,w] fp32 feature) is only used for the per-pixel norm.  Returns
 * SLOTVPS_EUNSUPPORTED when that head call did not run the tensor-core path for the finest level.
 * `workspace`: slotvps_mask_logits_workspace_bytes().                                              */
int slotvps_head_mask_logits(const slotvps_head_desc* d, void* head_workspace, size_t head_workspace_bytes, int frame,
                             const float* feat, const float* emb,
                             const float* feat_bn_w, const float* feat_bn_b, const float* feat_bn_mean, const float* feat_bn_var,
                             const float* fg_bn, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* BatchNorm2d in eval mode as a per-channel affine (feat_bn of generate_final_outputs, vps_temporal_slots.py:145-149):
 * scale[c] = w[c] / sqrt(var[c] + 1e-5), shift[c] = b[c] - mean[c] * scale[c].  Feeds slotvps_head_opts.feat_bn_scale/shift. */
int slotvps_fold_batchnorm(const float* w, const float* b, const float* mean, const float* var, int n,
                           float* scale, float* shift, void* stream);

/* Same; exactly one of `feat` (per-pixel norm computed from the fp32 feature) and `rnorm_ss` ([4][T][P_last] partial squared norms of
 * feat_bn(x) accumulated by slotvps_head_forward_ex with opts.rnorm_ss) is non-NULL.                                       */
int slotvps_head_mask_logits_ex(const slotvps_head_desc* d, void* head_workspace, size_t head_workspace_bytes, int frame,
                                const float* feat, const float* rnorm_ss, const float* emb,
                                const float* feat_bn_w, const float* feat_bn_b, const float* feat_bn_mean, const float* feat_bn_var,
                                const float* fg_bn, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Panoptic fusion: PostProcessPanopticInstances.forward (vps_temporal_slots.py:659-807, incl.
 * mask_removal :564-657) + the inline relabel of simple_test (:411-435), entirely on device.
 *   pred_logits [N,num_classes], pred_masks [N,h,w] (1/4 resolution), target size (H,W).
 * Outputs (device): panoptic [H,W] int64; meta int32[4 + 3*N]:
 *   meta[0]=K' kept slots, meta[1]=number of things, meta[2]=filter iterations, meta[3]=converged,
 *   then keep_index[N], label[N], (float bits) prob[N] in final order (stuff..., things...).
 *   masks_out (optional, may be NULL): [N_cap,H,W] fp32 masked logits of the kept slots
 *   (Instances.masks), only the first K' planes are written.                                 */
typedef struct slotvps_fusion_cfg {
  int32_t num_classes, stuff_num, small_area, max_iters;
  float threshold, pixel_threshold;   /* compared in fp32, as torch / numpy do (:691, :607); pixel_threshold must be > 1/3
                                         (the exact two-pass mask_removal keeps at most two candidates per pixel)          */
  double fraction_threshold;          /* compared in fp64, as numpy does (:621-622)              */
  int32_t logits_width;               /* columns of pred_logits: 0 or num_classes = with the "no object" column (:688-691),
                                         num_classes - 1 = without it, no class test (:692-693)                            */
  int32_t reserved;
} slotvps_fusion_cfg;
int slotvps_fusion_workspace_bytes(int n_slots, int H, int W, size_t* bytes);
int slotvps_panoptic_fuse(const slotvps_fusion_cfg* cfg, const float* pred_logits, const float* pred_masks,
                          int n_slots, int h, int w, int H, int W,
                          int64_t* panoptic, int32_t* meta, float* masks_out, int masks_cap,
                          void* workspace, size_t workspace_bytes, void* stream);
/* The reference's small-segment loop (:761-792) has no iteration bound; slotvps_panoptic_fuse launches cfg->max_iters
 * passes (each exits at once after the fixed point).  If meta[3] == 0 afterwards the id map holds the sentinel 255 and
 * this call runs `iters` more passes from the state left in `workspace` (same cfg / masks / shapes), then relabels.   */
int slotvps_panoptic_fuse_resume(const slotvps_fusion_cfg* cfg, const float* pred_masks, int n_slots, int h, int w, int H, int W,
                                 int64_t* panoptic, int32_t* meta, float* masks_out, int masks_cap,
                                 void* workspace, size_t workspace_bytes, int iters, void* stream);

/* Sine position embedding (PositionEmbeddingSine.forward, position_encoding.py:236-256,
 * normalize=True, 128 feats/axis): out [256,h,w].                                           */
int slotvps_sine_pos(float* out, int h, int w, void* stream);

/* ---- Tracker (first row beyond the retriever path) ------------------------------
 * SimpleTrackHead.forward (simple_track_head.py:58-92): x_query [k,256], ref_x_query [m,256] ->
 * match_score [k, 1+m] = [0 | fc(x_query) fc(ref_x_query)^T]; fc = `num_fcs` Linear(256,256) layers with ReLU
 * between them.  fc_w [num_fcs,256,256] ([out][in], as nn.Linear stores it), fc_b [num_fcs,256].
 * workspace: (k + m) * 256 floats.                                                                 */
int slotvps_track_scores(const float* fc_w, const float* fc_b, int num_fcs, const float* x_query, int k,
                         const float* ref_x_query, int m, float* match_score, void* workspace,
                         size_t workspace_bytes, void* stream);
/* The per-video tracking loop of simple_test (vps_temporal_slots.py:232-237 reset, :322-409 assignment) with the
 * object bank (prev_instances.output_embedding) kept in device memory.
 *   state: opaque device buffer of slotvps_track_state_bytes(capacity, n_slots); capacity = most objects per video.
 *   slotvps_track_reset: first frame of a video (fid == 1).
 *   slotvps_track_step:  embedding [N,256] = the head's last-stage slot embeddings of the current frame,
 *     fusion_meta = the meta written by slotvps_panoptic_fuse for that frame (kept slots, stuff first).
 *     track_out int32[4 + N] (device): [0]=K' entries, [1]=number of things, [2]=objects in the bank,
 *     [3]=1 if the bank overflowed `capacity` (ids stay valid, the extra embeddings are not remembered),
 *     then the object id of every kept entry; the last [1] of them are the reference's panoptic_det_obj_ids. */
int slotvps_track_state_bytes(int capacity, int n_slots, size_t* bytes);
int slotvps_track_reset(void* state, size_t state_bytes, void* stream);
int slotvps_track_step(const float* fc_w, const float* fc_b, int num_fcs, const float* embedding,
                       const int32_t* fusion_meta, int n_slots, void* state, size_t state_bytes, int capacity,
                       int32_t* track_out, void* stream);

/* ---- Consumers of the id map ----------------------------------------------------
 * Semantic prediction of simple_test (vps_temporal_slots.py:440-451): fcn_output [n_classes,h,w] fp32 -> bilinear
 * resize to (H,W) when the sizes differ (align_corners=False), softmax over classes, first index of the maximum.
 * out [H,W] int64 (the reference's fcn_outputs[0]).  n_classes <= 32.                                          */
int slotvps_semantic_argmax(const float* fcn_output, int n_classes, int h, int w, int H, int W, int64_t* out, void* stream);
/* CityscapesVps.get_unified_pan_result (tools/dataset/cityscapes_vps.py:214-302) for ONE frame, on device:
 *   seg [H,W] int64 semantic argmax, pan [H,W] int64 panoptic ids (< 256), cls_inds int32[n_inst] (thing class - 10),
 *   obj_ids int32[n_obj] or NULL / 0 (the reference's obj_ids=None), id_last_stuff = num_seg_classes - num_classes (:250),
 *   stuff_area_limit as passed by tools/test_vpq.py:169.  pan_2ch [H,W,3] uint8 = (semantic, instance from 1, object id + 1).
 * The counter for redundant object ids (:219 max_oid) lives in the workspace: slotvps_unify_reset() at the start of
 * what would be one get_unified_pan_result call, then one slotvps_unify_pan_result() per frame, in order.
 * status (optional, device int32[2]): [0] sticky error bits (1: id/class out of range, 2: cls_inds too short for an id
 * present in the map, 4: obj_ids too short -- the reference raises IndexError in the last two cases), [1] max_oid.  */
int slotvps_unify_workspace_bytes(size_t* bytes);
int slotvps_unify_reset(void* workspace, size_t workspace_bytes, void* stream);
int slotvps_unify_pan_result(const int64_t* seg, const int64_t* pan, const int32_t* cls_inds, int n_inst,
                             const int32_t* obj_ids, int n_obj, int H, int W, int id_last_stuff, int stuff_area_limit,
                             uint8_t* pan_2ch, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* ---- UPSNetFPN deformable-convolution subnet -------------------------------------
 * This is the one place on the widened path where the reference HAS a native boundary: the pybind'd
 * deform_conv_cuda.deform_conv_forward_cuda (mmdet/ops/dcn/src/deform_conv_cuda.cpp:152-245, bound at :688, called from
 * mmdet/ops/dcn/deform_conv.py:53-58).  slotvps_deform_conv_forward takes the same tensors and the same integer
 * arguments in the same order: input [B,c_in,H,W], weight [c_out,c_in,kH,kW], offset [B,2*kH*kW*deformable_group,H,W]
 * (channel 2*tap = dy, 2*tap+1 = dx, tap = ky*kW+kx; deform_conv_cuda_kernel.cu:216-226) -> output [B,c_out,H,W].
 * The reference's scratch arguments (`columns`, `ones`) become `workspace`; im2col_step only partitions the batch in the
 * reference and does not change the result, it is accepted and ignored.  Served here: what UPSNetFPN instantiates
 * (upsnetFPN.py:36-49) -- 3x3, stride 1, padding 1, dilation 1, group 1, deformable_group 1, c_in a multiple of 64 and
 * c_out a multiple of 32, both <= 256, B*H*W < 2^22 pixels; anything else returns SLOTVPS_EINVAL (the reference op stays available
 * for it).
 * offset == NULL computes the ordinary 3x3 convolution (all offsets zero).                                           */
int slotvps_deform_conv_workspace_bytes(int B, int c_in, int H, int W, size_t* bytes);
int slotvps_deform_conv_forward(const float* input, const float* weight, const float* offset, float* output,
                                int B, int c_in, int c_out, int H, int W,
                                int kW, int kH, int dW, int dH, int padW, int padH, int dilationW, int dilationH,
                                int group, int deformable_group, int im2col_step,
                                void* workspace, size_t workspace_bytes, void* stream);
/* The subnet UPSNetFPN applies to every FPN level with shared weights (upsnetFPN.py:36-49, 66-70):
 * n_layers x [DeformConvWithOffset 3x3 (mmdet/models/utils/deform_conv_with_offset.py: conv_offset = Conv2d(c_in,18,3,
 * padding=1) with bias, then the bias-free deformable conv) -> GroupNorm(32, c_out) -> ReLU].  Pointers are the module's
 * own parameters (state_dict keys deform_convs.0.{3i}.conv_offset.{weight,bias}, .{3i}.conv.weight, .{3i+1}.{weight,bias}).
 * x [B,c_in(0),H,W] -> out [B,c_out(last),H,W] (one entry of the reference's fpn_px list).  Activations stay
 * pixel-major between the layers; GroupNorm + ReLU of layer i are applied inside the loads of layer i+1.              */
typedef struct slotvps_dcn_layer {
  int32_t c_in, c_out;
  const float *offset_w, *offset_b;   /* conv_offset.weight [18,c_in,3,3], conv_offset.bias [18] */
  const float *weight;                /* conv.weight [c_out,c_in,3,3] */
  const float *gn_w, *gn_b;           /* GroupNorm(32, c_out) weight / bias [c_out] */
} slotvps_dcn_layer;
int slotvps_dcn_prepared_bytes(const slotvps_dcn_layer* layers, int n_layers, size_t* bytes);
int slotvps_dcn_prepare(const slotvps_dcn_layer* layers, int n_layers, void* prepared, size_t prepared_bytes, void* stream);
int slotvps_dcn_workspace_bytes(const slotvps_dcn_layer* layers, int n_layers, int B, int H, int W, size_t* bytes);
int slotvps_dcn_subnet_forward(const slotvps_dcn_layer* layers, int n_layers, const void* prepared, const float* x, float* out,
                               int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* Introspection. */
const char* slotvps_last_error(void);
const char* slotvps_version(void);
/* number of kernels launched by this library on the calling thread since the last reset */
int64_t slotvps_launch_count(int reset);
/* Per-launch device timing for the roofline report: between _begin and _end every kernel launch of
 * this thread records a CUDA event on its stream; _end synchronises them and writes rows
 * "kernel\tlaunches\ttotal_ms\n" (time between consecutive events) into buf.                  */
int slotvps_profile_begin(void* stream);
int slotvps_profile_end(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SLOTVPS_B200_H_ */
