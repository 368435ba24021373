/*
 * cpm_ops.h -- C ABI of the B200-native (sm_100a) CPM R-CNN detection-head op library (libcpm_ops.so).
 *
 * This is the drop-in boundary for the reference's native extension `pet.lib.ops._C`
 * (pybind module, /root/reference/pet/lib/ops/csrc/vision.cpp:21-22,32).  Each entry point cites the
 * reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; no torch / ATen types;
 *   - every pointer named `d_*` is a DEVICE pointer on the current CUDA device; the caller owns all
 *     memory (outputs and workspaces are allocated by the caller, e.g. with torch's caching allocator);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises
 *     the device (the reference does: ROIAlign_cuda.cu:422, ml_nms.cu:117);
 *   - return value: CPM_OK (0) or a negative CPM_ERR_* code; cpm_last_error() gives the message
 *     (thread-local).  The reference signals the same conditions with C++ exceptions
 *     (AT_ASSERTM / AT_ERROR -> Python RuntimeError); the Python host layer re-raises as RuntimeError;
 *   - there is no CPU fallback: a NULL/host pointer is an error, not a slow path.
 */
#ifndef CPM_OPS_H_
#define CPM_OPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CPM_API __attribute__((visibility("default")))
#else
#define CPM_API
#endif

#define CPM_OK 0
#define CPM_ERR_INVALID_ARG (-1)
#define CPM_ERR_UNSUPPORTED (-2)
#define CPM_ERR_WORKSPACE (-3)
#define CPM_ERR_CUDA (-4)

#define CPM_MAX_LEVELS 8

/* element types of feature maps / pooled outputs / gradients */
#define CPM_F32 0
#define CPM_F64 1   /* the reference dispatches float and double on the GPU (ROIAlign_cuda.cu:406) */
#define CPM_BF16 2  /* new on this path: bf16 storage, fp32 accumulation (NHWC kernels only) */

/* memory layout of one feature-map level */
#define CPM_LAYOUT_NCHW 0   /* (B, C, H, W) contiguous: what the reference's kernels index (ROIAlign_cuda.cu:219) */
#define CPM_LAYOUT_NHWC 1   /* (B, H, W, C) contiguous == a torch.channels_last (B,C,H,W) tensor; the fast kernels' layout */

/* interpolation_method of ROIAlign.h:57-65 (pet/lib/ops/roi_align.py:11) */
#define CPM_INTERP_BILINEAR 0
#define CPM_INTERP_NEAREST 1

/* how the fp32 union `Sa + Sb - inter` of the IoU is rounded (see oracle/cpm_oracle.c: orc_nms) */
#define CPM_IOU_PLAIN 0      /* every op rounded separately (source-level semantics, torchvision CPU)        */
#define CPM_IOU_TV_CUDA 1    /* fmaf(bw,bh,Sa) - inter : what torchvision 0.26's sm_100 nms kernel executes   */
#define CPM_IOU_ML_CUDA 2    /* fmaf(aw,ah,Sb) - inter : what nvcc 12.9 makes of the reference ml_nms.cu:23-25 */

/* backward accumulation mode */
#define CPM_BWD_DETERMINISTIC 0   /* atomic-free: pixel tiles own their gradient, fixed summation order */
#define CPM_BWD_ATOMIC 1          /* red.global.add.f32 scatter (non-deterministic order; reported separately) */

/* forward implementation selector (testing / benchmarking aid; AUTO is what callers use) */
#define CPM_FWD_AUTO 0
#define CPM_FWD_GENERIC 1         /* one thread per output element, any parameters / layout, fp32/fp64 */
#define CPM_FWD_NHWC 2            /* general channel-vector gather (bilinear, NHWC, fp32, C % 4 == 0, any pooled size) */
#define CPM_FWD_COLS 4            /* column-table kernel: 7x7 / 14x14 poolers, sampling_ratio 1|2, NHWC fp32 or bf16, either pooled layout */
#define CPM_FWD_ROWS 8            /* row-streaming TMA kernel: same poolers, fp32, (K,C,PH,PW) output, <= 4 levels (AUTO's choice where it
                                     applies; CPM_FWD_IMPL=cols in the environment makes AUTO take CPM_FWD_COLS instead) */

/*
 * A feature pyramid (or the pyramid of dense feature gradients): L levels of dense maps that share batch
 * size B, channel count C, element type and layout.  Replaces the list[Tensor] the reference's Pooler walks
 * (pet/rcnn/utils/poolers.py:103-132) and, with L == 1, the single `input` of
 * _C.roi_align_forward (ROIAlign.h:57-65).
 */
typedef struct cpm_pyramid {
  int32_t num_levels;                 /* L in [1, CPM_MAX_LEVELS] */
  int32_t batch;                      /* B */
  int32_t channels;                   /* C */
  int32_t dtype;                      /* CPM_F32 | CPM_F64 | CPM_BF16 */
  int32_t layout;                     /* CPM_LAYOUT_NCHW | CPM_LAYOUT_NHWC (all levels) */
  int32_t reserved;
  void* d_level[CPM_MAX_LEVELS];      /* level l: B*C*H_l*W_l elements, dense, in `layout` order */
  int32_t height[CPM_MAX_LEVELS];
  int32_t width[CPM_MAX_LEVELS];
  float spatial_scale[CPM_MAX_LEVELS];
} cpm_pyramid_t;

/*
 * FPN level heuristic, LevelMapper (poolers.py:9-40):
 *   lvl = clamp(floor(lvl0 + log2(sqrt(area) / s0 + eps)), k_min, k_max) - k_min,
 *   area = (x2 - x1 + 1) * (y2 - y1 + 1)                      (bounding_box.py:306-310).
 * Fused into the RoIAlign kernels when num_levels > 1.  `s / s0` is evaluated as s * (1 / s0), which is
 * how torch evaluates tensor / python-scalar on CUDA (the device the reference runs on).
 */
typedef struct cpm_level_mapper {
  float k_min, k_max;                 /* -log2(scales[0]), -log2(scales[-1])  (poolers.py:86-88) */
  float canonical_scale;              /* 224 */
  float canonical_level;              /* 4 */
  float eps;                          /* 1e-6 */
} cpm_level_mapper_t;

/* ---- library ------------------------------------------------------------------------------------ */
CPM_API const char* cpm_last_error(void);
CPM_API int cpm_version(void);
/* number of kernels this library has launched in this process (monotonic; bench.py's gpu_launches) */
CPM_API uint64_t cpm_launch_count(void);
/* binds the calling thread to CUDA device `device` for subsequent calls (CUDAGuard of ROIAlign_cuda.cu:383) */
CPM_API int cpm_set_device(int device);

/* ---- RoIAlign -------------------------------------------------------------------------------------
 * Forward.  Replaces _C.roi_align_forward (ROIAlign.h:57-96 -> ROIAlign_cuda.cu:367-425, kernel :178-256)
 * and, for num_levels > 1, the whole per-level loop of Pooler.forward (poolers.py:117-131): every RoI is
 * pooled from the level the mapper assigns it, results land at the RoI's own row of `d_out`
 * (K, C, PH, PW) -- one launch, no nonzero()/index_put, no device sync.
 *   d_rois      (K,5) [batch_idx, x1, y1, x2, y2], same dtype as the pyramid (ROIAlign_cuda.cu:381-382)
 *   mapper      ignored when num_levels == 1; d_roi_levels (optional, int32[K]) overrides the mapper
 *   impl        CPM_FWD_AUTO | CPM_FWD_GENERIC | CPM_FWD_NHWC | CPM_FWD_COLS | CPM_FWD_ROWS
 * K == 0 is a no-op (ROIAlign_cuda.cu:401-404). */
CPM_API int cpm_roi_align_forward(const cpm_pyramid_t* feat, const void* d_rois, int64_t K, int pooled_h, int pooled_w,
                          int sampling_ratio, int aligned, int interpolation, const cpm_level_mapper_t* mapper,
                          const int32_t* d_roi_levels, int impl, void* d_out, void* stream);

/* Backward.  Replaces _C.roi_align_backward (ROIAlign.h:98-146 -> ROIAlign_cuda.cu:428-487, kernel :259-365).
 * `grad_feat` describes the OUTPUT maps: every level is fully written (dense gradient, zeros where no RoI
 * reaches -- the reference's at::zeros + atomicAdd, :451-452,:340-347), also when K == 0.
 *   mode        CPM_BWD_DETERMINISTIC (needs the workspace) | CPM_BWD_ATOMIC
 * Workspace (deterministic mode): cpm_roi_align_backward_workspace_bytes(...) bytes, 256-byte aligned; it holds the
 * per-(level,image) RoI lists and the per-RoI tap tables (and, only for poolers with pooled size * sampling_ratio > 32,
 * which the single-pass staged kernel does not take, a channel-vector copy (K, PH*PW, C) of grad_out). */
CPM_API size_t cpm_roi_align_backward_workspace_bytes(int64_t K, int num_levels, int batch, int channels, int pooled_h,
                                                      int pooled_w, int sampling_ratio);
/* The same query given the gradient pyramid itself (shapes, dtype); what the Python layer calls.  Equal to the plain query
 * unless the experimental TMA tile kernel is selected (CPM_BWD_IMPL=tma: 7x7 / 14x14 poolers, sampling_ratio 2, C % 64 == 0,
 * fp32), whose per-tile stage lists it adds.  Both deterministic kernels write an NHWC or an NCHW gradient pyramid (the
 * layout _C.roi_align_backward returns, ROIAlign_cuda.cu:451-452). */
CPM_API size_t cpm_roi_align_backward_workspace_bytes_pyr(const cpm_pyramid_t* grad_feat, int64_t K, int pooled_h, int pooled_w,
                                                          int sampling_ratio);
CPM_API int cpm_roi_align_backward(const cpm_pyramid_t* grad_feat, const void* d_grad_out, const void* d_rois, int64_t K,
                           int pooled_h, int pooled_w, int sampling_ratio, int aligned, int interpolation,
                           const cpm_level_mapper_t* mapper, const int32_t* d_roi_levels, int mode,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* Memory layout of the POOLED tensor (forward output / backward grad_out).  The reference always produces
 * (K, C, PH, PW) contiguous (ROIAlign_cuda.cu:390-391); a head that runs in torch.channels_last (the conv stack of the
 * 14x14 grid head, grid_cascade_rcnn heads) exchanges the same logical tensor as (K, PH, PW, C), which saves both
 * kernels the (channel, bin) <-> channel-vector transposition.  The _ex entry points take the layout; the plain ones
 * above are the KCHW case.  KHWC is served by the fast kernels only (CPM_ERR_UNSUPPORTED otherwise: the caller converts). */
#define CPM_POOLED_KCHW 0
#define CPM_POOLED_KHWC 1
CPM_API int cpm_roi_align_forward_ex(const cpm_pyramid_t* feat, const void* d_rois, int64_t K, int pooled_h, int pooled_w,
                             int sampling_ratio, int aligned, int interpolation, const cpm_level_mapper_t* mapper,
                             const int32_t* d_roi_levels, int impl, int pooled_layout, void* d_out, void* stream);
CPM_API int cpm_roi_align_backward_ex(const cpm_pyramid_t* grad_feat, const void* d_grad_out, const void* d_rois, int64_t K,
                              int pooled_h, int pooled_w, int sampling_ratio, int aligned, int interpolation,
                              const cpm_level_mapper_t* mapper, const int32_t* d_roi_levels, int mode, int pooled_layout,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* LevelMapper alone (poolers.py:29-40): d_levels int64[K]. */
CPM_API int cpm_level_map(const float* d_rois, int64_t K, const cpm_level_mapper_t* mapper, int64_t* d_levels, void* stream);

/* ---- layout staging --------------------------------------------------------------------------------
 * One level (B, C, H, W) <-> (B, H, W, C).  `to_layout` is the layout of d_dst; d_src is in the other one.
 * The fast RoIAlign kernels read/write NHWC; torch.channels_last tensors need no staging at all. */
CPM_API int cpm_layout_convert(const void* d_src, void* d_dst, int batch, int channels, int height, int width, int dtype,
                       int to_layout, void* stream);
/* A whole pyramid in ONE launch (what Pooler.forward does for the NCHW maps the reference's FPN emits, FPN.py:96-121): `src`
 * and `dst` describe the same levels in the two layouts.  fp32 with 16-byte aligned levels takes a 16-byte-vectorised tile
 * transpose; anything else falls back to one cpm_layout_convert per level. */
CPM_API int cpm_layout_convert_pyramid(const cpm_pyramid_t* src, const cpm_pyramid_t* dst, void* stream);

/* ---- NMS ------------------------------------------------------------------------------------------
 * Hard NMS over N boxes (x1,y1,x2,y2), fp32.  Replaces
 *   d_labels == NULL : pet.lib.ops.nms = torchvision.ops.nms (pet/lib/ops/nms.py:2,10)
 *   d_labels != NULL : _C.ml_nms (ml_nms.h:16-39 -> ml_nms.cu:82-146): a pair is only compared when labels match.
 * Sort (stable, descending score), IoU bitmask and suppression sweep all run on the device.
 *   d_keep   int64[N]  indices into the input, descending score (ml_nms.cu:143-145); first *d_count are valid
 *   d_count  int64[1]  number kept; `topk` > 0 stops the sweep after topk keeps (ml_nms.cu:134)
 * N == 0 writes *d_count = 0. */
CPM_API size_t cpm_nms_workspace_bytes(int64_t N);
CPM_API int cpm_nms(const float* d_boxes, const float* d_scores, const int64_t* d_labels, int64_t N, float iou_threshold,
            int64_t topk, int iou_flavor, int64_t* d_keep, int64_t* d_count, void* d_workspace,
            size_t workspace_bytes, void* stream);

/* Batched NMS: independent segments (image x class, or image x FPN level) in ONE call -- the per-image /
 * per-level Python loops of rpn/inference.py:102-113 and grid_cascade_rcnn/inference.py:91-97.
 *   d_segments int32[N] segment id of each box in [0, num_segments)
 *   d_keep     int64[N] kept indices grouped by segment (ascending id), descending score inside a segment
 *   d_seg_counts int64[num_segments] kept per segment (after topk_per_segment, 0 = unlimited)
 *   d_count    int64[1] total kept */
CPM_API size_t cpm_nms_batched_workspace_bytes(int64_t N, int64_t num_segments);
CPM_API int cpm_nms_batched(const float* d_boxes, const float* d_scores, const int32_t* d_segments, int64_t N,
                    int64_t num_segments, float iou_threshold, int64_t topk_per_segment, int iou_flavor,
                    int64_t* d_keep, int64_t* d_seg_counts, int64_t* d_count, void* d_workspace,
                    size_t workspace_bytes, void* stream);

/* ---- grid-point decode ------------------------------------------------------------------------------
 * Replaces GridPostProcessor.get_boxes (grid_cascade_rcnn/inference.py:189-279), which moves the heat-maps
 * to the host.  d_logits (R,P,h,w) fp32 PRE-sigmoid; d_boxes (R,4); sub_xy = HOST int32[P*2] (sub_x1, sub_y1)
 * of calc_sub_regions (grid_rcnn/loss.py:244-273); P must be a square number <= 64.
 *   d_out_boxes (R,4) un-clamped (the reference's clamp_ at :275-276 is a no-op)
 *   d_out_scores (R,P) max sigmoid per point, may be NULL */
CPM_API int cpm_grid_decode(const float* d_logits, const float* d_boxes, int64_t R, int P, int h, int w,
                    const int32_t* sub_xy, float mapping_ratio, float* d_out_boxes, float* d_out_scores,
                    void* stream);

/* ---- RPN proposal decode (next row) ---------------------------------------------
 * For all (FPN level, image) candidate sets of a batch in one launch: BoxCoder.decode (box_coder.py:51-94),
 * clip_to_image (bounding_box.py:294-304) and remove_small_boxes (boxlist_ops.py:104-118), the per-image body of
 * RPNPostProcessor.forward_for_single_feature_map (rpn/inference.py:96-113) up to the NMS.
 *   d_deltas, d_anchors (M,4) fp32; d_segments int32[M] in [0, num_segments); d_segment_im_wh (num_segments,2) = (width,
 *   height) of each segment's image; weights = HOST float[4] (wx, wy, ww, wh); bbox_xform_clip = log(1000/16).
 *   d_boxes (M,4) decoded + clipped; d_segments_out int32[M]: the input segment, or one of the num_trash trash segments
 *   num_segments .. num_segments + num_trash - 1 for a box smaller than min_size (spread round-robin so that no trash
 *   segment outgrows the NMS's per-segment limit) -- feed both to cpm_nms_batched with num_segments + num_trash segments. */
CPM_API int cpm_rpn_decode(const float* d_deltas, const float* d_anchors, const int32_t* d_segments,
                   const float* d_segment_im_wh, int64_t M, int64_t num_segments, int64_t num_trash,
                   const float* weights, float bbox_xform_clip, float min_size, float* d_boxes, int32_t* d_segments_out,
                   void* stream);

/* All FPN levels of RPNPostProcessor.forward (rpn/inference.py:115-143) in two launches around ONE top-k:
 *   cpm_rpn_flatten_objectness: rows (L*N, row) fp32, row l*N+n = permute_and_flatten (utils/misc.py:6-10) of level l's
 *     objectness logits (N, A, H, W) for image n, i.e. index hw*A + a, padded with -inf up to `row` -- the caller applies
 *     sigmoid and one top-k (k = max_l k[l], sorted) to the rows (inference.py:84-88 for every level at once);
 *   cpm_rpn_select_decode: for level l, image n, rank j < k[l], in that order (output index first[l] + n*k[l] + j with
 *     first[l] = N * sum_{l' < l} k[l']): the winner's deltas from the (N, 4A, H, W) regression output, its anchor
 *     (d_anchors[l] is (H*W*A, 4), or (N, H*W*A, 4) when anchors_per_image[l]), then cpm_rpn_decode's arithmetic against
 *     d_image_wh (N,2); d_scores = the top-k values; segments = l*N + n or a trash segment (too small, or a padding winner).
 *   M = N * sum_l k[l] outputs; feed boxes / scores / segments to cpm_nms_batched with L*N + num_trash segments. */
typedef struct {
  int32_t num_levels, num_images;
  int32_t row;                                  /* padded row length, >= max_l A[l]*HW[l] */
  int32_t reserved;
  const float* d_objectness[CPM_MAX_LEVELS];    /* (N, A, H, W) fp32 contiguous */
  const float* d_regression[CPM_MAX_LEVELS];    /* (N, 4A, H, W) fp32 contiguous */
  const float* d_anchors[CPM_MAX_LEVELS];
  int32_t anchors_per_image[CPM_MAX_LEVELS];
  int32_t A[CPM_MAX_LEVELS], HW[CPM_MAX_LEVELS], k[CPM_MAX_LEVELS];
} cpm_rpn_levels_t;
CPM_API int cpm_rpn_flatten_objectness(const cpm_rpn_levels_t* levels, float* d_rows, void* stream);
CPM_API int cpm_rpn_select_decode(const cpm_rpn_levels_t* levels, const int64_t* d_topk_idx, const float* d_topk_val,
                          int64_t top_k, const float* d_image_wh, int64_t num_trash, const float* weights,
                          float bbox_xform_clip, float min_size, float* d_boxes, float* d_scores, int32_t* d_segments_out,
                          void* stream);

/* ---- grid-point training targets (next row) ------------------------------------
 * Replaces GridLossComputation.prepare_target (grid_cascade_rcnn/loss.py:178-258: CPU triple loop + H2D).
 *   d_pos_boxes, d_gt_boxes (R,4) fp32: positive RoIs and their matched ground truth; sub_xy = HOST int32[P*2] sub-region
 *   offsets (calc_sub_regions, grid_rcnn/loss.py:244-273); map_size = 4 * roi_feat_size (56);
 *   d_targets (R, P, map_size/2, map_size/2) fp32 of 0/1, fully written. */
CPM_API int cpm_grid_targets(const float* d_pos_boxes, const float* d_gt_boxes, int64_t R, int grid_points, int map_size,
                     const int32_t* sub_xy, float mapping_ratio, int pos_radius, int target_refine, float* d_targets,
                     void* stream);

/* ---- box IoU matrix + Matcher (next row) ----------------------------------------
 * cpm_box_iou: boxlist_iou (pet/utils/data/structures/boxlist_ops.py:123-158): d_iou (N,M) = IoU with the +1 area
 *   convention of boxes1 (N,4) against boxes2 (M,4), fp32, the reference's operation order.
 * cpm_matcher: Matcher.__call__ (pet/rcnn/utils/matcher.py:52-112) on an (M,N) quality matrix (M ground truth x N
 *   predictions): d_matches int64[N] = index of the best ground truth, -1 below low_threshold, -2 between the thresholds;
 *   allow_low_quality_matches restores every prediction that ties some ground truth's best overlap.  No host sync. */
CPM_API int cpm_box_iou(const float* d_boxes1, const float* d_boxes2, int64_t N, int64_t M, float* d_iou, void* stream);
CPM_API size_t cpm_matcher_workspace_bytes(int64_t M, int64_t N);
CPM_API int cpm_matcher(const float* d_quality, int64_t M, int64_t N, float high_threshold, float low_threshold,
                int allow_low_quality_matches, int64_t* d_matches, void* d_workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CPM_OPS_H_ */
