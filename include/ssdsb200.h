/*
 * ssdsb200.h — C ABI of libssdsb200.so, the B200 (sm_100a) single-shot-detector hot path.
 *
 * This is the drop-in boundary for the `ssds._C` extension module that the reference
 * (ShuangXieIrene/ssds.pytorch @ b5ec682) names but never shipped:
 *   ssds/modeling/layers/box.py:3-4      commented imports  `from ssds._C import decode, nms`
 *   ssds/modeling/layers/box.py:419-421  call site  decode_cuda(cls_head, box_head, anchors, stride, threshold, top_n)
 *   ssds/modeling/layers/box.py:483-485  call site  nms_cuda(scores, boxes, classes, nms, ndetections)
 *   ssds/utils/export.py:134-139         `hasattr(ssds, "_C")` feature probe
 * plus the operator surface the same path uses in pure python: generate_anchors (box.py:46-58),
 * box2delta/delta2box (:61-87), extract_targets/snap_to_anchors_by_iou (:116-226, :362-405),
 * MultiBoxLoss.forward (ssds/core/criterion.py:43-71) and the conv stack of
 * ssds/modeling/ssds/ssd.py:42-74 over ssds/modeling/nets/resnet.py:41-56.
 *
 * Conventions
 *  - Plain C: raw DEVICE pointers + sizes + a cudaStream_t (passed as void*); no torch types.
 *  - The caller owns every buffer, including the workspace (`*_workspace_bytes` tells how much);
 *    the library never allocates or frees device memory and keeps no mutable global state.
 *  - Every output is fully written (zero padded exactly like the reference), so no memset is needed.
 *  - No host synchronisation; all work is enqueued on `stream`; fixed launch shapes
 *    (CUDA-graph capturable).
 *  - Return value: 0 on success, negative ssdsb_status otherwise; never throws across the ABI.
 *    `ssdsb_last_error_string()` gives the detail for the calling thread.
 *  - All float tensors are contiguous fp32 unless a name says bf16; index outputs are int32.
 */
#ifndef SSDSB200_H_
#define SSDSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDSB_API __attribute__((visibility("default")))

typedef enum {
  SSDSB_OK = 0,
  SSDSB_ERR_INVALID_ARGUMENT = -1, /* maps to ValueError in the python shim            */
  SSDSB_ERR_WORKSPACE = -2,        /* workspace too small / NULL                        */
  SSDSB_ERR_CUDA = -3,             /* a CUDA runtime/driver call or a launch failed     */
  SSDSB_ERR_UNSUPPORTED = -4       /* valid request outside what the kernels implement  */
} ssdsb_status;

#define SSDSB_MAX_LEVELS 8

#define SSDSB_ABI_VERSION 202

SSDSB_API int ssdsb_version(void);          /* == SSDSB_ABI_VERSION of the header the library was built from */

/* ABI handshake for bindings (ctypes / cgo / JNI): out4 = {SSDSB_ABI_VERSION, sizeof(ssdsb_conv_desc),
 * sizeof(ssdsb_level), SSDSB_MAX_LEVELS}.  A binding compares these with its own mirror of the structs and
 * refuses to load a library built from a different header. */
SSDSB_API int ssdsb_abi_info(int* out4);
SSDSB_API const char* ssdsb_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * Anchors ("PriorBox").  reference: box.py:46-58 generate_anchors, grid add box.py:151-159.
 * ------------------------------------------------------------------------------------------- */
/* Base anchors [A,4], A = n_ratios*n_scales, scale-major / ratio-minor, round-half-even. */
SSDSB_API int ssdsb_generate_anchors(int stride, const float* h_ratios, int n_ratios,
                                     const float* h_scales, int n_scales,
                                     float* d_out /*[A,4]*/, void* stream);
/* Materialised grid in the reference's order [A, W, H, 4] (x-major, box.py:151-159). */
SSDSB_API int ssdsb_anchor_grid(const float* d_base /*[A,4]*/, int A, int stride, int W, int H,
                                float* d_out /*[A,W,H,4]*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Box codec.  reference: box2delta box.py:61-71 ("encode"), delta2box box.py:74-87 ("decode").
 * ------------------------------------------------------------------------------------------- */
SSDSB_API int ssdsb_box2delta(const float* d_boxes, const float* d_anchors, int n,
                              float* d_out /*[n,4]*/, void* stream);
SSDSB_API int ssdsb_delta2box(const float* d_deltas, const float* d_anchors, int n, int size_w,
                              int size_h, int stride, float* d_out /*[n,4]*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * decode: threshold + top-k + index->(a,c,y,x) + loc gather + delta2box + clamp + centerness
 * rescore.  reference: box.py:408-477 (per level) and decoder.py:36-48 (all levels, concatenated
 * along dim 1 in level order, each level padded to top_n).
 * Ties between equal scores are broken by ascending flat index (the reference leaves it to
 * torch.topk).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* conf;    /* [B, A*C, H, W] scores (already sigmoid-ed), channel = a*C + c */
  const float* loc;     /* [B, A*4, H, W] deltas, channel = a*4 + k                      */
  const float* anchors; /* [A, 4] base anchors (device)                                  */
  int A, C, H, W, stride;
} ssdsb_level;

SSDSB_API size_t ssdsb_decode_workspace_bytes(const ssdsb_level* levels, int n_levels, int B,
                                              int top_n);
SSDSB_API int ssdsb_decode(const ssdsb_level* levels, int n_levels, int B, float threshold,
                           int top_n, int rescore,
                           float* d_scores /*[B, L*top_n]*/, float* d_boxes /*[B, L*top_n, 4]*/,
                           float* d_classes /*[B, L*top_n]*/,
                           int32_t* d_index /*[B, L*top_n] flat index in the level, -1 pad; may be NULL*/,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * nms: batched, class-aware (D)IoU NMS with the reference's exact arithmetic
 * (+1 pixel areas, 1e-7 eps, DIoU on top-left corners).  reference: box.py:480-546.
 * ------------------------------------------------------------------------------------------- */
SSDSB_API size_t ssdsb_nms_workspace_bytes(int B, int N, int ndetections);
SSDSB_API int ssdsb_nms(const float* d_scores /*[B,N]*/, const float* d_boxes /*[B,N,4]*/,
                        const float* d_classes /*[B,N]*/, int B, int N, float nms_threshold,
                        int ndetections, int using_diou,
                        float* d_out_scores /*[B,D]*/, float* d_out_boxes /*[B,D,4]*/,
                        float* d_out_classes /*[B,D]*/,
                        int32_t* d_out_index /*[B,D] position in the input row, -1 pad; may be NULL*/,
                        float* d_out_packed /*[B,D,6] (score, x1, y1, x2, y2, class) zero padded, the block
                                              ssds.py:60-68 returns / the ranks all-gather; may be NULL.  When
                                              given, the three separate outputs may all be NULL.*/,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * match: IoU arg-max matcher + encode + depth (+ one-hot class target).
 * reference: extract_targets box.py:362-405 -> snap_to_anchors_by_iou box.py:116-226.
 * targets [B,T,5] = (x, y, w, h, label), rows with label <= -1 are padding.
 * d_cls_target may be NULL (fused consumers only need depth).
 * ------------------------------------------------------------------------------------------- */
SSDSB_API int ssdsb_match_iou(const float* d_targets /*[B,T,5]*/, int B, int T,
                              const float* d_base_anchors /*[A,4]*/, int A, int C, int stride,
                              int H, int W, float match_threshold, float unmatch_threshold,
                              float center_sampling_radius,
                              float* d_cls_target /*[B,A,C,H,W] or NULL*/,
                              float* d_box_target /*[B,A,4,H,W]*/, float* d_depth /*[B,A,1,H,W]*/,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * MultiBoxLoss: BCE-with-logits + per-image hard-negative mining.
 * reference: ssds/core/criterion.py:43-71 (intended per-image semantics).
 *  - ssdsb_multibox_loss      drop-in: unreduced loss [B,A,C,H,W] from logits + one-hot target.
 *  - ssdsb_multibox_loss_sum  fused: no materialised target; the class comes from depth
 *    (depth > 0 => class = depth-1); writes per image sum(loss * (depth >= 0)) and the number of
 *    positives, i.e. what pipeline_anchor_basic.py:76-82 reduces to.
 * ------------------------------------------------------------------------------------------- */
SSDSB_API size_t ssdsb_multibox_loss_workspace_bytes(int B, int A, int C, int H, int W);
SSDSB_API int ssdsb_multibox_loss(const float* d_logits /*[B,A,C,H,W]*/,
                                  const float* d_target /*[B,A,C,H,W]*/,
                                  const float* d_depth /*[B,A,1,H,W]*/, int B, int A, int C, int H,
                                  int W, int negpos_ratio, float* d_out /*[B,A,C,H,W]*/,
                                  void* d_workspace, size_t workspace_bytes, void* stream);
SSDSB_API int ssdsb_multibox_loss_sum(const float* d_logits /*[B,A,C,H,W]*/,
                                      const float* d_depth /*[B,A,1,H,W]*/, int B, int A, int C,
                                      int H, int W, int negpos_ratio, float* d_loss_sum /*[B]*/,
                                      float* d_num_pos /*[B]*/, void* d_workspace,
                                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Focal / SmoothL1 / IoU-family losses, forward.  reference: ssds/core/criterion.py
 * FocalLoss :95-108, SmoothL1Loss :138-151, IOULoss :175-239; masking / summation as the caller does
 * (pipeline_anchor_basic.py:76-97).
 *  - ssdsb_focal_loss / ssdsb_loc_loss        drop-in, unreduced outputs.
 *  - ssdsb_focal_loss_sum / ssdsb_loc_loss_sum fused: per image sum(loss * (depth >= 0)) (cls, class
 *    taken from depth) and sum(loss * (depth > 0)) (loc); + number of positives.
 * loc `type`: 0 SmoothL1 (beta), 1 iou, 2 giou, 3 diou, 4 ciou.  pred/target deltas [B,A,4,H,W];
 * unreduced loc output is [B,A,4,H,W] for SmoothL1 and [B,A,1,H,W] for the IoU family.
 * ------------------------------------------------------------------------------------------- */
SSDSB_API size_t ssdsb_loss_sum_workspace_bytes(int B, int A, int H, int W);
SSDSB_API int ssdsb_focal_loss(const float* d_logits, const float* d_target, int B, int A, int C, int H,
                               int W, float alpha, float gamma, float* d_out, void* stream);
SSDSB_API int ssdsb_focal_loss_sum(const float* d_logits, const float* d_depth, int B, int A, int C, int H,
                                   int W, float alpha, float gamma, float* d_loss_sum /*[B]*/,
                                   float* d_num_pos /*[B]*/, void* d_workspace, size_t workspace_bytes,
                                   void* stream);
SSDSB_API int ssdsb_loc_loss(const float* d_pred, const float* d_target, int B, int A, int H, int W,
                             int type, float beta, float* d_out, void* stream);
SSDSB_API int ssdsb_loc_loss_sum(const float* d_pred, const float* d_target, const float* d_depth, int B,
                                 int A, int H, int W, int type, float beta, float* d_loss_sum /*[B]*/,
                                 void* d_workspace, size_t workspace_bytes, void* stream);

/* Backward of the fused per-image sums ("plus backward so the training step is
 * usable"): gradients of  sum_b scale[b] * loss_sum[b]  w.r.t. the logits / predicted deltas, where
 * loss_sum is what ssdsb_{multibox,focal,loc}_loss_sum return; the caller folds 1/fg_targets and the
 * upstream gradient into d_scale [B].  Same values autograd gives on the reference modules
 * (criterion.py:43-239 under pipeline_anchor_basic.py:76-97); the mined hard-negative mask and ciou's
 * alpha are constants there too.  Outputs have the shape of the differentiated input. */
SSDSB_API int ssdsb_multibox_loss_sum_backward(const float* d_logits, const float* d_depth, int B, int A,
                                               int C, int H, int W, int negpos_ratio, const float* d_scale,
                                               float* d_grad_logits, void* d_workspace,
                                               size_t workspace_bytes, void* stream);
SSDSB_API int ssdsb_focal_loss_sum_backward(const float* d_logits, const float* d_depth, int B, int A, int C,
                                            int H, int W, float alpha, float gamma, const float* d_scale,
                                            float* d_grad_logits, void* stream);
SSDSB_API int ssdsb_loc_loss_sum_backward(const float* d_pred, const float* d_target, const float* d_depth,
                                          int B, int A, int H, int W, int type, float beta,
                                          const float* d_scale, float* d_grad_pred, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The loss half of one training step in ONE launch (loss_step.cu): for every level and image,
 * extract_targets (IoU matcher, box.py:362-405) + the classification criterion (MultiBoxLoss with per-image
 * hard-negative mining, criterion.py:43-71, or FocalLoss :95-108) * (depth >= 0) + the localisation criterion
 * (SmoothL1 :138-151 or iou/giou/diou/ciou :175-239) * (depth > 0), summed, and the caller's normalisation
 *   fg_targets = sum_l max(#(depth_l > 0), 1);  cls_loss = sum / fg_targets;  loc_loss = sum / fg_targets
 * exactly as ssds/pipeline/pipeline_anchor_basic.py:62-97 sequences them.  Reads every logit once; no one-hot
 * target, and no depth / box_target tensor unless the level asks for them.  Deterministic; no host sync;
 * CUDA-graph capturable.  MATCHER.CENTER_SAMPLING_RADIUS must be 0 (the default): the class of a positive is
 * depth-1.  BCE uses ex2.approx + a degree-8 log1p polynomial (relative error <= 5e-7 per element; hard
 * negatives whose max-CE differ by less than that may swap at the cut).
 * ------------------------------------------------------------------------------------------- */
typedef struct ssdsb_loss_level {
  const float* conf;    /* [B, A*C, H, W] raw logits (model in training mode, ssd.py:72-73)            */
  const float* loc;     /* [B, A*4, H, W] raw deltas; may be NULL when loc_kind == SSDSB_LOC_NONE      */
  const float* anchors; /* [A, 4] base anchors (device, 16-byte aligned)                                */
  float* depth;         /* optional out [B,A,1,H,W] (what extract_targets returns), or NULL            */
  float* box_target;    /* optional out [B,A,4,H,W], or NULL                                            */
  int A, C, H, W, stride;
} ssdsb_loss_level;

enum { SSDSB_CLS_MULTIBOX = 0, SSDSB_CLS_FOCAL = 1 };
enum { SSDSB_LOC_NONE = -1, SSDSB_LOC_SMOOTHL1 = 0, SSDSB_LOC_IOU = 1, SSDSB_LOC_GIOU = 2, SSDSB_LOC_DIOU = 3,
       SSDSB_LOC_CIOU = 4 };

SSDSB_API size_t ssdsb_detection_loss_workspace_bytes(const ssdsb_loss_level* levels, int n_levels, int B);
SSDSB_API int ssdsb_detection_loss(const ssdsb_loss_level* levels, int n_levels, int B,
                                   const float* d_targets /*[B,T,5] (x,y,w,h,label), label <= -1: padding*/, int T,
                                   float match_threshold, float unmatch_threshold,
                                   int cls_kind, int negpos_ratio, float focal_alpha, float focal_gamma,
                                   int loc_kind, float smoothl1_beta,
                                   float* d_out_scalars /*[3] cls_loss, loc_loss, fg_targets*/,
                                   float* d_out_cls_sum /*[L*B] per (level, image), may be NULL*/,
                                   float* d_out_loc_sum /*[L*B], may be NULL*/,
                                   float* d_out_num_pos /*[L*B], may be NULL*/,
                                   void* d_workspace /*256-byte aligned*/, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conv stack (tcgen05 / TMA implicit GEMM), bf16 x bf16 -> fp32 accumulate.
 * One call replaces nn.Conv2d -> BatchNorm2d(eval, folded) -> [+ residual] -> [ReLU] of the
 * reference model graph (ssds/modeling/ssds/ssd.py:42-74, nets/resnet.py:41-56, torchvision
 * Bottleneck, layers/basic_layers.py:41-57), or a multibox head pair (ssd.py:100-103) + the eval
 * sigmoid (ssd.py:72-73).
 *   x : NHWC bf16, channel stride x_cstride (0 => Cin); Cin % 16 == 0 (K-blocks of 64, 32 or 16
 *       channels = 128/64/32-byte swizzle rows; Cin == 16 is also the space-to-depth packed image of
 *       the stride-2 stems, see ssdsb_pack_image_s2d).
 *   w : bf16 [w_rows >= Cout][KH*KW][Cin], BN scale folded in; bias fp32 [Cout] (folded BN shift
 *       or the conv bias).
 *   out_mode 0 : y = NHWC bf16 [N,Ho,Wo,out_cstride], optional residual (same layout, res_cstride)
 *                and activation (relu: 0 none, 1 ReLU, 2 ReLU6).  Cout % 32 == 0.
 *   out_mode 1 : multibox head: output channels [0,n_loc) -> y  = fp32 NCHW [N,n_loc,Ho,Wo] (loc),
 *                [n_loc,Cout) -> y2 = fp32 NCHW [N,Cout-n_loc,Ho,Wo] (conf), sigmoid-ed if `sigmoid`.
 *   Ho/Wo: 0 => (H + 2*pad - KH)/stride + 1; pad is the top/left padding, the bottom/right halo is
 *   whatever Ho/Wo imply (TMA zero-fills out-of-bounds reads).
 * ------------------------------------------------------------------------------------------- */
enum { SSDSB_CONV_OUT_NHWC_BF16 = 0, SSDSB_CONV_OUT_HEAD_NCHW_F32 = 1 };
/* x_kind: 0 = plain NHWC.  1 = "windowed stem": x is the space-to-depth image of
 * ssdsb_pack_image_s2d written with 2 zero pixels of left padding into rows of x_row_pixels
 * (>= W + 3) pixels; KH = KW = 4, Cin = 16; one K-block = the 4 horizontally adjacent taps x 16
 * channels = 64 contiguous bf16, fetched through an overlapping-window tensor map. */
enum { SSDSB_CONV_X_PLAIN = 0, SSDSB_CONV_X_WINDOWED_STEM = 1 };

typedef struct {
  int N, H, W, Cin;
  int Cout, KH, KW, stride, pad;
  int Ho, Wo;
  int x_cstride, out_cstride, res_cstride;
  int x_row_pixels; /* pixels per image row in memory (0 => W) */
  int chunk;        /* 0 = dense.  > 0: block-diagonal ("grouped") conv, Cin == Cout, every `chunk`
                     * consecutive channels form an independent conv (chunk % 32 == 0, <= 128; a
                     * grouped conv with group width gw uses chunk = lcm(gw, 32)).  w is then
                     * bf16 [(Cout/chunk)*128][KH*KW][chunk]: one 128-row slab per chunk, rows >= chunk
                     * zero.  NHWC output only. */
  int x_kind;
  int w_rows;
  int relu;
  int out_mode;
  int n_loc;
  int sigmoid;
} ssdsb_conv_desc;

SSDSB_API int ssdsb_conv2d_bf16(const ssdsb_conv_desc* desc, const void* d_x, const void* d_w,
                                const float* d_bias, const void* d_residual, void* d_y, void* d_y2,
                                void* stream);

/* Introspection (no reference counterpart; used by the parity tests and bench.py's self-check to PROVE
 * which instantiation a shape ran): fills out8 with what the calling thread's last ssdsb_conv2d_bf16
 * launched: {BLOCK_N, BLOCK_K, ways (interleaved accumulators), weight-resident flag, pipeline stages,
 * grid, groups, m_tiles % ways (!= 0: the last group runs ghost tiles)}. */
SSDSB_API int ssdsb_conv_last_launch(int* out8);

/* Two chained pointwise convolutions in one launch (conv_pair.cu):
 *   y1 = act1(conv1x1(x, w1) + bias1 [+ residual]);   y2 = act2(conv1x1(y1, w2) + bias2)
 * i.e. a torchvision Bottleneck's conv3+bn3+add+ReLU followed by the NEXT block's conv1+bn1+ReLU
 * (reference nets/resnet.py:41-56).  Layer 2 reads each y1 tile back from L2 right after the same CTA
 * stored it, so y1 is written once and never re-read from HBM.  x [N,H,W,Cin], y1/residual
 * [N,H,W,Cmid], y2 [N,H,W,Cout2], all dense NHWC bf16; w1 [Cmid,Cin], w2 [Cout2,Cmid] bf16 (BN folded);
 * channel counts multiples of 64 (above 256: of 256); relu: 0 none, 1 ReLU, 2 ReLU6.  Results are
 * bit-identical to two ssdsb_conv2d_bf16 calls. */
SSDSB_API int ssdsb_conv1x1_pair_bf16(int N, int H, int W, int Cin, int Cmid, int Cout2, int relu1, int relu2,
                                      const void* d_x, const void* d_w1, const float* d_bias1,
                                      const void* d_residual, void* d_y1, const void* d_w2,
                                      const float* d_bias2, void* d_y2, void* stream);

/* Fused MobileNetV2 inverted residual (conv_mbconv.cu; reference ssds/modeling/nets/mobilenet.py:40-76, the
 * torchvision InvertedResidual it builds its backbone from):
 *   h = act_e(conv1x1(x, w_exp) + b_exp)            [skipped when d_w_exp == NULL: h = x, hid == Cin]
 *   g = act_d(depthwise3x3(h, w_dw, stride, pad 1) + b_dw)
 *   y = act_p(conv1x1(g, w_proj) + b_proj) [+ x]
 * in ONE launch: the hid-channel tensors h and g only ever exist tile by tile in shared memory / TMEM.
 * x [N,H,W,Cin], y [N,Ho,Wo,Cout] NHWC bf16 (Ho = (H-1)/stride + 1); w_exp [w_exp_rows >= hid, Cin],
 * w_proj [w_proj_rows >= Cout, hid] bf16 K-major (BN folded), w_dw [9, hid] bf16 (ssdsb_dwconv3x3 layout),
 * biases fp32.  Channel counts are multiples of 32, Cout <= 256 (SSDSB_ERR_UNSUPPORTED otherwise: use the three
 * separate launches).  Results are bit-identical to ssdsb_conv2d_bf16 -> ssdsb_dwconv3x3_nhwc_bf16 ->
 * ssdsb_conv2d_bf16 (same bf16 rounding points, same fp32 accumulation order). */
typedef struct {
  int N, H, W;
  int Cin, hid, Cout;
  int stride;       /* of the depthwise conv: 1 | 2 */
  int residual;     /* 1: y += x (needs stride 1 and Cin == Cout) */
  int relu_expand, relu_dw, relu_project; /* 0 none, 1 ReLU, 2 ReLU6 */
  int w_exp_rows, w_proj_rows;
  int out_cstride;  /* channel stride of y in elements; 0 => Cout */
} ssdsb_mbconv_desc;

SSDSB_API int ssdsb_mbconv_bf16(const ssdsb_mbconv_desc* desc, const void* d_x, const void* d_w_exp,
                                const float* d_b_exp, const void* d_w_dw, const float* d_b_dw,
                                const void* d_w_proj, const float* d_b_proj, void* d_y, void* stream);
/* Introspection for tests / tools: what the calling thread's last ssdsb_mbconv_bf16 launched:
 * {hidden chunk, tile W, tile H, expand M tiles per chunk, chunks per tile, x buffers, staging slots,
 *  depthwise row segments, rows per segment, grid, dynamic shared memory bytes, project accumulator columns}. */
SSDSB_API int ssdsb_mbconv_last_launch(int* out12);
/* Diagnostic (host-synchronising, not for the hot path): with SSDSB_MB_PROF set in the environment, CTA 0 of every
 * ssdsb_mbconv_bf16 launch records how many cycles each of its 16 warps spent blocked on each barrier id
 * (out[warp * 16 + id], ids listed in conv_mbconv.cu), out[256] = cycles CTA 0 ran, out[257] = chunks it processed,
 * out[258 + (g - 8) * 8 + k] = cycle stamp of event k of chunk g = 8..15 (dw start / dw done / project wait / project
 * issue / expand committed / convert start / convert done).  The buffer holds 322 entries. */
SSDSB_API int ssdsb_mbconv_profile(unsigned long long* out322);

/* Image pre-processing fused with the layout change the stem needs (SSDDetector.__call__,
 * ssds/ssds.py:48-57: HWC->CHW, (x - mean)/std): packs an image batch into the 2x2
 * space-to-depth NHWC16 bf16 tensor [N, H/2, W/2, 16] (channel (a*2+b)*3+c = pixel (2i+a, 2j+b),
 * channels 12..15 zero) on which the 7x7/s2 stem is a 4x4/s1 convolution.
 * src_format 0: fp32 NCHW [N,3,H,W]; 1: uint8 NHWC [N,H,W,3].  H and W must be even. */
SSDSB_API int ssdsb_pack_image_s2d(const void* d_src, int src_format, int N, int H, int W,
                                   float mean, float std, int out_row_pixels /*0 => W/2*/,
                                   int left_pad /*pixels skipped at the start of each output row;
                                   the caller zeroes the padding once*/,
                                   void* d_out, void* stream);

/* 3x3 / stride 2 / pad 1 max pooling on NHWC bf16 (resnet.py:45 `self.maxpool`). C % 8 == 0. */
SSDSB_API int ssdsb_maxpool3x3s2_nhwc_bf16(const void* d_x, int N, int H, int W, int C, void* d_y,
                                           void* stream);
/* 5x5 / stride 1 / pad 2 max-pool, NHWC bf16, with channel strides (elements) for input and output so that it can
 * read one channel slice of a concatenated buffer and write another: YOLOv4's SPP block (reference
 * ssds/modeling/ssds/yolo.py:161-184, pools of 5 / 9 / 13 concatenated after x) is three cascaded calls. */
SSDSB_API int ssdsb_maxpool5x5s1_nhwc_bf16(const void* d_x, int N, int H, int W, int C, int x_cstride, void* d_y,
                                           int y_cstride, void* stream);

/* Depthwise 3x3 / pad 1 / stride 1|2 conv + folded BN + activation (relu: 0 none, 1 ReLU, 2 ReLU6) on
 * NHWC bf16 — torchvision InvertedResidual's depthwise stage (reference nets/mobilenet.py:78) and
 * SepConvBNReLU (layers/basic_layers.py:5-24).  w: bf16 [9][C] (tap-major), bias fp32 [C], C % 8 == 0. */
SSDSB_API int ssdsb_dwconv3x3_nhwc_bf16(const void* d_x, const void* d_w, const float* d_bias, int N,
                                        int H, int W, int C, int stride, int relu, void* d_y,
                                        void* stream);

/* BiFPN weighted fusion (ssds/modeling/ssds/bifpn.py:41-62), NHWC bf16:
 *   mode 0: out = w0*a + w1*nearest_up2(b)            a,out [N,H,W,C], b [N,H/2,W/2,C]
 *   mode 1: out = w0*a + w1*maxpool2x2(b) [+ w2*c]    a,c,out [N,H,W,C], b [N,2H,2W,C]; c may be NULL
 * w* are the relu-normalised scalars (host-side, bifpn.py:35-38). */
SSDSB_API int ssdsb_bifpn_fuse_nhwc_bf16(const void* d_a, const void* d_b, const void* d_c, int mode,
                                         float w0, float w1, float w2, int N, int H, int W, int C,
                                         void* d_out, void* stream);

/* FPN top-down merge (ssds/modeling/ssds/fpn.py:80-87): fine[n,h,w,:] += coarse[n,h/2,w/2,:]
 * (nearest 2x upsample + add), NHWC bf16, in place on `fine` ([N,H,W,C]; coarse is [N,H/2,W/2,C]). */
SSDSB_API int ssdsb_upsample2x_add_nhwc_bf16(const void* d_coarse, void* d_fine, int N, int H, int W,
                                             int C, void* stream);

/* YOLOv3 top-down merge (reference ssds/modeling/ssds/yolo.py:70-72): out = cat(fine, nearest_up2(coarse)) along
 * channels; fine [N,H,W,Cf], coarse [N,H/2,W/2,Cc], out [N,H,W,Cf+Cc], dense NHWC bf16, Cf % 8 == Cc % 8 == 0. */
SSDSB_API int ssdsb_upsample2x_concat_nhwc_bf16(const void* d_fine, const void* d_coarse, int N, int H, int W,
                                                int Cf, int Cc, void* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSDSB200_H_ */
