/* libmxdet_sm100.so - C ABI of the B200-native detection hot path.
 *
 * Drop-in boundary for the data-parallel hot path of jiangzhengkai/mxdetection
 * (SURVEY.md section 8b).  The mounted reference holds no source, so each
 * entry cites the README line of the module it serves and the MXNet 1.3 /
 * mmdetection-0.5 operator whose contract it implements.
 *
 * Conventions
 *  - Tensors cross as `const DLTensor*` (borrowed; see mxdet_dlpack.h).  They
 *    must live on ONE CUDA device (kDLCUDA / kDLCUDAManaged), be compact
 *    row-major (strides NULL or canonical) and have exactly the dtype stated.
 *    A CPU tensor is refused with MXD_ENOTSUP: there is no CPU fallback.
 *  - The caller owns every input, output and workspace.  The library never
 *    allocates, frees or synchronises; every call only enqueues kernels on
 *    `stream` (a cudaStream_t passed as void*, NULL = legacy default stream)
 *    and is CUDA-graph capturable.
 *  - Outputs of data-dependent length have fixed capacity plus a device
 *    int32 count.
 *  - Return value 0 (MXD_OK) or a negative MXD_E* code; the message is in
 *    the thread-local mxd_last_error().  Re-entrant and callable from several
 *    host threads: the only global state is the thread-local error string, an
 *    atomic launch counter and the once-per-device kernel attribute flags
 *    (dynamic shared-memory opt-in), which are set under a mutex on first use.
 *  - delta ("legacy +1"): 1.0 for the py-faster-rcnn / mmdet-0.5 lineage
 *    (widths x2-x1+1), 0.0 for mx.nd.contrib.box_nms / box_iou.
 *  - All IoU / threshold / label arithmetic is fp32 round-to-nearest without
 *    FMA contraction, in the order of SURVEY.md 8(a) Specs A-H.
 */
#ifndef MXDET_H_
#define MXDET_H_
#include <stddef.h>
#include <stdint.h>
#include "mxdet_dlpack.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MXD_OK 0
#define MXD_EINVAL (-1)     /* bad shape / dtype / stride / attribute            */
#define MXD_ENOTSUP (-2)    /* CPU tensor, or size outside the supported range   */
#define MXD_ECUDA (-3)      /* CUDA runtime error at enqueue time                */
#define MXD_EWORKSPACE (-4) /* workspace NULL or too small                       */

#define MXD_MAX_LEVELS 8    /* FPN levels per call                               */
#define MXD_MAX_BASE_ANCHORS 16
#define MXD_SORT_CAP 8192   /* rows sorted inside one CTA; longer top-k / NMS take the chunked path */

/* ---- library ------------------------------------------------------------ */
int mxd_version(void);                /* 10000*major + 100*minor + patch        */
const char* mxd_last_error(void);     /* thread-local, valid until next call    */
uint64_t mxd_launch_count(void);      /* kernels enqueued by this library so far */
/* Host plumbing of the host-buffer RoI stage (ops.HostRoIStage): strided copy of `rows` runs of
 * `width_bytes` between pinned host memory and the device (kind 1 = H2D, 2 = D2H), async on stream. */
int mxd_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                     size_t width_bytes, size_t rows, int kind, void* stream);

/* ---- A1/A2  RoIAlign  (mxdetection/ops, /root/reference/README.md:24;
 *      contract mx.nd.contrib.ROIAlign + _backward_ROIAlign, mxnet 1.3.0
 *      /root/reference/README.md:37; Spec A) -------------------------------- */
/* data (N,C,H,W) f32; rois (R,5) f32 [b,x1,y1,x2,y2]; out (R,C,PH,PW) f32.     */
/* Workspace for the plane-resident kernels (shared-memory resident row bands of
 * the feature planes).  workspace == NULL selects the gather
 * kernels (same results, slower).  feat_h/feat_w: host int[num_levels].          */
size_t mxd_roi_align_workspace_bytes(int num_rois, int batch, int channels, int num_levels,
                                     const int* feat_h, const int* feat_w,
                                     int pooled_h, int pooled_w, int sample_ratio);
int mxd_roi_align_forward(const DLTensor* data, const DLTensor* rois, DLTensor* out,
                          int pooled_h, int pooled_w, float spatial_scale, int sample_ratio,
                          void* workspace, size_t workspace_bytes, void* stream);
/* grad_out (R,C,PH,PW); grad_data (N,C,H,W).  accumulate=0: req='write'
 * (grad_data is zero-filled first); accumulate=1: req='add'.  grad wrt rois is
 * identically zero and not produced.                                            */
int mxd_roi_align_backward(const DLTensor* grad_out, const DLTensor* rois, DLTensor* grad_data,
                           int pooled_h, int pooled_w, float spatial_scale, int sample_ratio,
                           int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- G1/G2  FPN level assignment + multi-level RoI extraction
 *      (mxdetection/models/roi_extractors, /root/reference/README.md:32;
 *      SingleLevelRoI.map_roi_levels / .forward of mmdet 0.5; Spec G) -------- */
/* rois (R,5) or (R,4) f32 -> levels (R) i32.                                   */
int mxd_map_roi_levels(const DLTensor* rois, DLTensor* levels, int num_levels,
                       float finest_scale, void* stream);
/* feats[l] (N,C,H_l,W_l) f32; levels (R) i32 or NULL (then Spec G is evaluated
 * in-kernel with finest_scale); spatial_scales host float[num_levels].          */
int mxd_roi_align_fpn_forward(const DLTensor* const* feats, int num_levels,
                              const float* spatial_scales, const DLTensor* rois,
                              const DLTensor* levels, DLTensor* out,
                              int pooled_h, int pooled_w, int sample_ratio, float finest_scale,
                              void* workspace, size_t workspace_bytes, void* stream);
int mxd_roi_align_fpn_backward(const DLTensor* grad_out, const DLTensor* rois,
                               const DLTensor* levels, DLTensor* const* grad_feats,
                               int num_levels, const float* spatial_scales,
                               int pooled_h, int pooled_w, int sample_ratio, float finest_scale,
                               int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- B1  NMS  (mxdetection/ops, /root/reference/README.md:24; contract
 *      mx.nd.contrib.box_nms, Spec B: strict iou > thr, stable score-desc /
 *      index-asc order) ---------------------------------------------------- */
/* Stable top-k: scores (S,n) f32 -> idx (S,k) i32, vals (S,k) f32 sorted by
 * (score desc, index asc); k = min(topk,n) (topk<=0: k=n).  k <= MXD_SORT_CAP
 * is sorted inside one CTA / cluster; larger k takes a chunk-sort + rank-merge
 * path that needs the workspace mxd_topk_stable_workspace_bytes reports.        */
size_t mxd_topk_stable_workspace_bytes(int segments, int n, int topk);
int mxd_topk_stable(const DLTensor* scores, DLTensor* idx, DLTensor* vals, int topk,
                    void* workspace, size_t workspace_bytes, void* stream);
/* boxes (n,4), scores (n) f32; ids (n) i32 or NULL; keep (cap) i32 receives
 * original row indices in score order, num_keep (1) i32 the count
 * (min(kept,cap)).  Rows with score <= valid_thresh are dropped first (pass
 * -INFINITY to keep all); topk<=0 = all; max_out<=0 = unlimited.
 * force_suppress=0 suppresses only rows with equal ids.                         */
size_t mxd_nms_workspace_bytes(int n, int topk);
int mxd_nms(const DLTensor* boxes, const DLTensor* scores, const DLTensor* ids,
            DLTensor* keep, DLTensor* num_keep, float iou_thr, float delta, int topk,
            float valid_thresh, int force_suppress, int max_out,
            void* workspace, size_t workspace_bytes, void* stream);
/* Batched form over ragged segments of ONE (n,4) / (n) pair (e.g. the (image, level) segments of the proposal
 * stage): segment s = rows [seg_offsets[s], seg_offsets[s+1]); seg_offsets (S+1) i32 ascending, on the device;
 * max_seg_len = host upper bound of the longest segment (sizes the workspace).  keep (S,cap) i32 receives GLOBAL row
 * indices in score order, -1 padded; num_keep (S) i32.  Other arguments as mxd_nms, applied per segment.            */
size_t mxd_nms_batched_workspace_bytes(int num_segments, int max_seg_len, int topk);
int mxd_nms_batched(const DLTensor* boxes, const DLTensor* scores, const DLTensor* ids,
                    const DLTensor* seg_offsets, int max_seg_len, DLTensor* keep, DLTensor* num_keep,
                    float iou_thr, float delta, int topk, float valid_thresh, int force_suppress,
                    int max_out, void* workspace, size_t workspace_bytes, void* stream);
/* MXNet tensor form: data (B,N,K) f32 -> out (B,N,K) kept rows first (score
 * order), all other rows -1; index (B,N) i32 or NULL receives the source row of
 * each output row (-1 for padding) - the record box_nms' backward consumes.
 * in_format/out_format: 0 corner, 1 center.                                      */
size_t mxd_box_nms_workspace_bytes(int batch, int n, int topk);
int mxd_box_nms(const DLTensor* data, DLTensor* out, DLTensor* index, float overlap_thresh,
                float valid_thresh, int topk, int coord_start, int score_index, int id_index,
                int force_suppress, int in_format, int out_format,
                void* workspace, size_t workspace_bytes, void* stream);
/* _backward_box_nms: in_grad[b,index[b,r],:] = out_grad[b,r,:], zero elsewhere. */
int mxd_box_nms_backward(const DLTensor* out_grad, const DLTensor* index, DLTensor* in_grad,
                         void* stream);

/* ---- C1/C2  anchors  (mxdetection/core/anchor, /root/reference/README.md:16;
 *      AnchorGenerator.grid_anchors / valid_flags, anchor_inside_flags of
 *      mmdet 0.5; Spec C) -------------------------------------------------- */
/* base_anchors: host float[num_base*4]; out (feat_h*feat_w*num_base, 4) f32,
 * order (y,x,a).                                                               */
int mxd_grid_anchors(const float* base_anchors, int num_base, int feat_h, int feat_w,
                     float stride, DLTensor* out, void* stream);
/* flags (feat_h*feat_w*num_base) u8: cell valid iff y<valid_h && x<valid_w.     */
int mxd_valid_flags(int feat_h, int feat_w, int valid_h, int valid_w, int num_base,
                    DLTensor* flags, void* stream);
/* out = valid & x1>=-ab & y1>=-ab & x2<img_w+ab & y2<img_h+ab (ab<0: = valid).  */
int mxd_inside_flags(const DLTensor* anchors, const DLTensor* valid, int img_h, int img_w,
                     float allowed_border, DLTensor* out, void* stream);

/* ---- D1/D2  IoU + max-IoU assigner  (mxdetection/core/bbox and core/anchor,
 *      /root/reference/README.md:16-17; bbox_overlaps /
 *      bbox_assign_wrt_overlaps of mmdet 0.5, mx.nd.contrib.box_iou; Specs D,E) */
/* b1 (G,4), b2 (N,4) -> out (G,N) f32 (materialising; tests and small G*N).     */
int mxd_bbox_overlaps(const DLTensor* b1, const DLTensor* b2, DLTensor* out, float delta,
                      void* stream);
/* Fused (never materialises G x N).  anchors (N,4); gts (B,G,4) [or (G,4), B=1];
 * num_gts (B) i32 or NULL (= G each); gt_labels (B,G) i32 or NULL; flags (N) or
 * (B,N) u8 or NULL.  Outputs (B,N): assigned i32 (-1 ignore, 0 neg, g+1 pos),
 * max_overlaps f32, labels i32 (or NULL).                                        */
size_t mxd_max_iou_assign_workspace_bytes(int batch, int num_gt);
int mxd_max_iou_assign(const DLTensor* anchors, const DLTensor* gts, const DLTensor* num_gts,
                       const DLTensor* gt_labels, const DLTensor* flags,
                       DLTensor* assigned, DLTensor* max_overlaps, DLTensor* labels,
                       float pos_iou_thr, float neg_iou_thr, float min_pos_iou, float delta,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- N2  detection post-processing  (SURVEY.md 8(f) "next" row N2: mxdetection/models/bbox_heads,
 *      /root/reference/README.md:29; BBoxHead.get_det_bboxes / multiclass_nms of mmdet 0.5).
 *      boxes (n,4) or (n,4*C); deltas NULL (boxes are final) or (n,4) / (n,4*C) decoded per Spec F on (n,4) boxes,
 *      clipped to (img_h,img_w) when both > 0 and divided by scale_factor; cls_score (n,C) softmaxed, column 0 =
 *      background.  Candidates (box i, class c >= 1) with score > score_thr enter ONE class-aware NMS (ids = c-1,
 *      at most MXD_SORT_CAP of them, best first); dets (cap,5) [x1,y1,x2,y2,score] / labels (cap) i32 (c-1), rows
 *      past num (1) i32 are 0 / -1; cap = min(min(n*(C-1), MXD_SORT_CAP), max_per_img) (max_per_img <= 0: no cap). */
size_t mxd_det_bboxes_workspace_bytes(long long n, int num_classes, int max_per_img);
int mxd_det_bboxes(const DLTensor* boxes, const DLTensor* deltas, const DLTensor* cls_score, const float* means,
                   const float* stds, int img_h, int img_w, double wh_ratio_clip, float scale_factor, float score_thr,
                   float iou_thr, float delta, int max_per_img, DLTensor* dets, DLTensor* labels, DLTensor* num,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- N1  random sampling + target packing  (SURVEY.md 8(f) "next" row N1: mxdetection/core/anchor + core/bbox,
 *      /root/reference/README.md:16-17; RandomSampler / anchor_target_single / bbox_target_single of mmdet 0.5).
 *      RNG contract: keys (N) f32 in [0,1) are supplied by the caller; the sample is the positives (assigned > 0) /
 *      negatives (assigned == 0) with the LARGEST keys, ties to the lower index.  kp = min(int(num*pos_fraction), N), the product taken in double
 *      (pos_fraction is a double so that the host side's int(num * fraction) and the library agree for 0.7, 0.9, ...)
 *      positives at most, negatives fill up to num (at most neg_pos_ub * max(1, num_pos) when neg_pos_ub >= 0).
 *      pos_inds (kp) / neg_inds (min(num, N)) i32, -1 padded; counts (2) i32 = {num_pos, num_neg}.            */
size_t mxd_random_sample_workspace_bytes(long long n, int num);
int mxd_random_sample(const DLTensor* assigned, const DLTensor* keys, int num, double pos_fraction, int neg_pos_ub,
                      DLTensor* pos_inds, DLTensor* neg_inds, DLTensor* counts, void* workspace,
                      size_t workspace_bytes, void* stream);
/* labels (N) i32, label_weights (N) f32, bbox_targets (N,4), bbox_weights (N,4): zero everywhere except the sampled
 * rows - positives: label 1 (or gt_labels[g]), weight pos_weight (<= 0: 1), Spec F deltas to their GT, box weight 1;
 * negatives: label weight 1.  gt_labels (G) i32 or NULL.                                                      */
int mxd_pack_targets(const DLTensor* anchors, const DLTensor* assigned, const DLTensor* gts, const DLTensor* gt_labels,
                     const DLTensor* pos_inds, const DLTensor* neg_inds, const float* means, const float* stds,
                     float pos_weight, DLTensor* labels, DLTensor* label_weights, DLTensor* bbox_targets,
                     DLTensor* bbox_weights, void* stream);

/* ---- N4  mask targets / mask paste  (SURVEY.md 8(f) "next" row N4: mxdetection/core/mask and models/mask_heads,
 *      /root/reference/README.md:18,30; mask_target / FCNMaskHead.get_seg_masks of mmdet 0.5).
 *      Target: gt_masks (G,H,W) u8, proposals (P,>=4) f32 image coords, gt_inds (P) i32 -> target (P,S,S), u8
 *      [RoIAlign(mask, roi, S, scale 1, sample_ratio) >= thr] or f32 (the RoIAlign values); Spec A in strict fp32 on
 *      the mask bytes, bit-exact.  Paste: mask_pred (n,C,S,S) or (n,S,S) f32 probabilities, labels (n) i32 or NULL
 *      (class c reads channel c+1), det_bboxes (n,>=4) -> im_masks (n,img_h,img_w) u8; box = trunc(bbox /
 *      scale_factor), half-pixel bilinear resize of the S x S map to the box, > thr (Spec N4).            */
int mxd_mask_target(const DLTensor* gt_masks, const DLTensor* proposals, const DLTensor* gt_inds,
                    DLTensor* target, int mask_size, int sample_ratio, float thr, void* stream);
int mxd_paste_masks(const DLTensor* mask_pred, const DLTensor* labels, const DLTensor* det_bboxes,
                    DLTensor* im_masks, float scale_factor, float thr, void* stream);

/* ---- F1/F2  delta encode / decode + clip  (mxdetection/core/bbox,
 *      /root/reference/README.md:17; bbox2delta / delta2bbox of mmdet 0.5;
 *      Spec F).  means/stds: host float[4].  exp/log correctly rounded fp32.   */
int mxd_bbox2delta(const DLTensor* proposals, const DLTensor* gts, DLTensor* deltas,
                   const float* means, const float* stds, void* stream);
/* max_h/max_w <= 0: no clipping.                                               */
int mxd_delta2bbox(const DLTensor* rois, const DLTensor* deltas, DLTensor* boxes,
                   const float* means, const float* stds, int max_h, int max_w,
                   double wh_ratio_clip, void* stream);

/* ---- H1  RPN proposal stage  (mxdetection/models/rpn_heads,
 *      /root/reference/README.md:28; RPNHead.get_proposals of mmdet 0.5 /
 *      mx.nd.contrib.MultiProposal; Spec H) -------------------------------- */
typedef struct {
  int num_levels;
  int feat_h[MXD_MAX_LEVELS];
  int feat_w[MXD_MAX_LEVELS];
  float stride[MXD_MAX_LEVELS];
  int num_base;                                         /* A, same on every level */
  float base_anchors[MXD_MAX_LEVELS][MXD_MAX_BASE_ANCHORS][4];
  int nms_pre;          /* per-level pre-NMS top-k (<=0: all)                    */
  int nms_post;         /* per-level post-NMS cap                                */
  int max_num;          /* per-image output rows                                 */
  float nms_thr;
  float min_bbox_size;  /* <=0: no filter                                        */
  float means[4];
  float stds[4];
  float delta;          /* 1.0                                                   */
  double wh_ratio_clip; /* 16/1000 (double: |log| is rounded to fp32 once)       */
} mxd_rpn_config;
/* scores[l] (B, H_l*W_l*A) f32 activated, (y,x,a) order; deltas[l]
 * (B, H_l*W_l*A, 4) f32; img_shapes (B,2) i32 [h,w] on device;
 * proposals (B,max_num,5) f32 [x1,y1,x2,y2,score] zero padded; num_valid (B) i32. */
size_t mxd_rpn_proposals_workspace_bytes(const mxd_rpn_config* cfg, int batch);
/* kmax = max_l min(nms_pre, n_l) (row stride of the per-segment stage buffers),
 * keep_stride = min(nms_post, kmax).                                            */
/* sizeof(mxd_rpn_config) as compiled into the library (binding self-check).       */
int mxd_sizeof_rpn_config(void);
int mxd_rpn_proposals_dims(const mxd_rpn_config* cfg, int* kmax, int* keep_stride);
int mxd_rpn_proposals(const DLTensor* const* scores, const DLTensor* const* deltas,
                      const DLTensor* img_shapes, const mxd_rpn_config* cfg,
                      DLTensor* proposals, DLTensor* num_valid,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Stage-wise outputs of the last mxd_rpn_proposals call that used `workspace`
 * (for the stage-wise parity tests): copies, per (image b, level l) segment,
 * the sorted top-k indices / decoded boxes / keep positions into caller
 * buffers.  idx (B,L,kmax) i32 (-1 padded); boxes (B,L,kmax,4) f32;
 * keep (B,L,keep_stride) i32 positions into the sorted rows (-1 padded);
 * counts (B,L,2) i32 [n_sorted,n_keep].                                         */
int mxd_rpn_proposals_stages(const mxd_rpn_config* cfg, int batch, const void* workspace,
                             size_t workspace_bytes, DLTensor* idx, DLTensor* boxes,
                             DLTensor* keep, DLTensor* counts, void* stream);

/* Tail of the data-parallel path (the detections every GPU hands to the final gather): packs
 * proposals (B,M,5) [x1,y1,x2,y2,score] + num_valid (B) into ONE buffer packed (B,M+1,6) f32 that a single
 * all-gather moves - row 0 of image b = {count, image id, 0,0,0,0}, rows 1..M = {image id (-1: padding row),
 * x1,y1,x2,y2,score}; image id = first_image_id + b.                                                       */
int mxd_pack_detections(const DLTensor* proposals, const DLTensor* num_valid, int first_image_id,
                        DLTensor* packed, void* stream);

/* mx.nd.contrib.MultiProposal / Proposal of mxnet 1.3.0 (multi_proposal.cc/.cu; SURVEY.md 8(a) Spec H
 * alt-mode; the call site would be mxdetection/models/rpn_heads, /root/reference/README.md:28).
 * cls_prob (N,2A,H,W) f32 [foreground = channels A..2A-1], bbox_pred (N,4A,H,W), im_info (N,3)
 * [height,width,scale] -> rois (N*post_n,5) [batch,x1,y1,x2,y2] and, when not NULL, scores (N*post_n,1);
 * rows beyond the kept boxes repeat them cyclically as mxnet does.  base_anchors: host float[A*4]
 * (utils::GenerateAnchors table).  rpn_pre_nms_top_n <= 0 = all.                                     */
size_t mxd_multi_proposal_workspace_bytes(int batch, int num_anchors, int feat_h, int feat_w,
                                          int rpn_pre_nms_top_n, int rpn_post_nms_top_n);
int mxd_multi_proposal(const DLTensor* cls_prob, const DLTensor* bbox_pred, const DLTensor* im_info,
                       DLTensor* rois, DLTensor* scores, const float* base_anchors, int num_anchors,
                       float feature_stride, int rpn_pre_nms_top_n, int rpn_post_nms_top_n,
                       float threshold, float rpn_min_size, void* workspace, size_t workspace_bytes,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MXDET_H_ */
