// RPN proposal decode for the batched proposal selection.
//
// Replaces, for ALL (FPN level, image) candidate sets of a batch at once, the per-level / per-image chain of
//   BoxCoder.decode            pet/rcnn/utils/box_coder.py:51-94
//   BoxList.clip_to_image      pet/utils/data/structures/bounding_box.py:294-304  (remove_empty=False)
//   remove_small_boxes         pet/utils/data/structures/boxlist_ops.py:104-118
// of RPNPostProcessor.forward_for_single_feature_map (pet/rcnn/modeling/rpn/inference.py:96-113): one thread per
// candidate, same operation order as the torch expressions (the library is compiled without FMA contraction), boxes that
// fail the size test are moved to trash segments instead of being compacted away (no host sync), so the batched NMS that
// follows never sees them.
#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

struct DecodeParams {
  float inv_w[4];          // applied as a division, like rel_codes[:, k::4] / w_k
  float clip, min_size;
};

__global__ void __launch_bounds__(256) rpn_decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                                                          const int* __restrict__ seg_in, const float2* __restrict__ seg_wh,
                                                          long M, int num_segments, int num_trash, float wx, float wy, float ww, float wh,
                                                          float clip, float min_size, float4* __restrict__ boxes,
                                                          int* __restrict__ seg_out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float4 a = anchors[i], d = deltas[i];
  const int seg = seg_in[i];
  const float widths = (a.z - a.x) + 1.0f, heights = (a.w - a.y) + 1.0f;              // TO_REMOVE = 1, box_coder.py:63-65
  const float ctr_x = a.x + 0.5f * widths, ctr_y = a.y + 0.5f * heights;
  const float dx = d.x / wx, dy = d.y / wy;
  const float dw = fminf(d.z / ww, clip), dh = fminf(d.w / wh, clip);                 // torch.clamp(max=bbox_xform_clip)
  const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;                  // mul, then add (separate torch ops)
  const float pw = expf(dw) * widths, ph = expf(dh) * heights;
  float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph;
  float x2 = (pcx + 0.5f * pw) - 1.0f, y2 = (pcy + 0.5f * ph) - 1.0f;
  int so = num_segments + (int)(i % num_trash);                                       // a trash segment
  if (seg >= 0 && seg < num_segments) {
    const float2 im = seg_wh[seg];                                                    // (width, height) of the segment's image
    x1 = fminf(fmaxf(x1, 0.0f), im.x - 1.0f);                                         // clamp_(min=0, max=size - 1)
    y1 = fminf(fmaxf(y1, 0.0f), im.y - 1.0f);
    x2 = fminf(fmaxf(x2, 0.0f), im.x - 1.0f);
    y2 = fminf(fmaxf(y2, 0.0f), im.y - 1.0f);
    const float ws = (x2 - x1) + 1.0f, hs = (y2 - y1) + 1.0f;                         // xywh conversion of remove_small_boxes
    if (ws >= min_size && hs >= min_size) so = seg;
  }
  boxes[i] = make_float4(x1, y1, x2, y2);
  seg_out[i] = so;
}

// ---- all levels in two launches: objectness rows for ONE top-k, and gather + decode of its winners ------------------------
struct RpnLevelsView {
  int num_levels, num_images, row;     // row = length of a padded objectness row (>= the largest H*W*A)
  const float* obj[CPM_MAX_LEVELS];
  const float* reg[CPM_MAX_LEVELS];
  const float* anchors[CPM_MAX_LEVELS];
  int per_image[CPM_MAX_LEVELS];
  int A[CPM_MAX_LEVELS], HW[CPM_MAX_LEVELS], k[CPM_MAX_LEVELS];
  long first[CPM_MAX_LEVELS + 1];      // output offset of the level: sum of N * k over the levels before it
};

// rows[(l * N + n) * row + hw * A + a] = objectness[l][n][a][hw] (permute_and_flatten, pet/rcnn/utils/misc.py:6-10, for every
// level and image at once), -inf behind the level's H*W*A entries
__global__ void __launch_bounds__(256) rpn_flatten_kernel(RpnLevelsView lv, float* __restrict__ rows) {
  const int r = blockIdx.y;                       // l * N + n
  const int l = r / lv.num_images, n = r - l * lv.num_images;
  const int A = lv.A[l], HW = lv.HW[l];
  const float* src = lv.obj[l] + (long)n * A * HW;
  float* dst = rows + (long)r * lv.row;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lv.row; i += gridDim.x * blockDim.x) {
    float v = -INFINITY;
    if (i < A * HW) {
      const int hw = i / A, a = i - hw * A;
      v = __ldg(src + (long)a * HW + hw);
    }
    dst[i] = v;
  }
}

// one thread per winner of the top-k: its regression deltas from the (N, 4A, H, W) head output, its anchor, then exactly
// rpn_decode_kernel's arithmetic; output order = level, image, rank (what the per-level loop concatenates)
__global__ void __launch_bounds__(256) rpn_select_decode_kernel(RpnLevelsView lv, const long long* __restrict__ top_idx,
                                                                 const float* __restrict__ top_val, int top_k,
                                                                 const float2* __restrict__ image_wh, long M, int num_trash,
                                                                 float wx, float wy, float ww, float wh, float clip,
                                                                 float min_size, float4* __restrict__ boxes,
                                                                 float* __restrict__ scores, int* __restrict__ seg_out) {
  const long m = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (m >= M) return;
  int l = 0;
  while (l + 1 < lv.num_levels && m >= lv.first[l + 1]) l++;
  const long e = m - lv.first[l];
  const int k = lv.k[l], N = lv.num_images;
  const int n = (int)(e / k), j = (int)(e - (long)n * k);
  const int A = lv.A[l], HW = lv.HW[l];
  const long src = ((long)(l * N + n)) * top_k + j;
  const long long i = top_idx[src];
  const int num_segments = lv.num_levels * N;
  int so = num_segments + (int)(m % num_trash);
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  scores[m] = top_val[src];
  if (i >= 0 && i < (long long)A * HW) {
    const int hw = (int)(i / A), a = (int)(i - (long long)hw * A);
    const float* rg = lv.reg[l] + ((long)n * 4 * A + 4 * a) * HW + hw;
    const float4 d = make_float4(__ldg(rg), __ldg(rg + HW), __ldg(rg + 2 * (long)HW), __ldg(rg + 3 * (long)HW));
    const float4 an = __ldg(reinterpret_cast<const float4*>(lv.anchors[l]) + (lv.per_image[l] ? (long)n * A * HW : 0L) + i);
    const float widths = (an.z - an.x) + 1.0f, heights = (an.w - an.y) + 1.0f;
    const float ctr_x = an.x + 0.5f * widths, ctr_y = an.y + 0.5f * heights;
    const float dx = d.x / wx, dy = d.y / wy;
    const float dw = fminf(d.z / ww, clip), dh = fminf(d.w / wh, clip);
    const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
    const float pw = expf(dw) * widths, ph = expf(dh) * heights;
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph;
    float x2 = (pcx + 0.5f * pw) - 1.0f, y2 = (pcy + 0.5f * ph) - 1.0f;
    const float2 im = image_wh[n];
    x1 = fminf(fmaxf(x1, 0.0f), im.x - 1.0f);
    y1 = fminf(fmaxf(y1, 0.0f), im.y - 1.0f);
    x2 = fminf(fmaxf(x2, 0.0f), im.x - 1.0f);
    y2 = fminf(fmaxf(y2, 0.0f), im.y - 1.0f);
    const float ws = (x2 - x1) + 1.0f, hs = (y2 - y1) + 1.0f;
    if (ws >= min_size && hs >= min_size) so = l * N + n;
    bx = make_float4(x1, y1, x2, y2);
  }
  boxes[m] = bx;
  seg_out[m] = so;
}

static int make_levels_view(const cpm_rpn_levels_t* lv, RpnLevelsView& v, bool need_obj, bool need_reg) {
  CPM_CHECK_ARG(lv != nullptr, "levels is NULL");
  CPM_CHECK_ARG(lv->num_levels >= 1 && lv->num_levels <= CPM_MAX_LEVELS && lv->num_images >= 1, "bad level / image count");
  v.num_levels = lv->num_levels;
  v.num_images = lv->num_images;
  v.row = lv->row;
  long first = 0;
  for (int l = 0; l < CPM_MAX_LEVELS; l++) {
    const bool on = l < lv->num_levels;
    v.obj[l] = on ? lv->d_objectness[l] : nullptr;
    v.reg[l] = on ? lv->d_regression[l] : nullptr;
    v.anchors[l] = on ? lv->d_anchors[l] : nullptr;
    v.per_image[l] = on ? lv->anchors_per_image[l] : 0;
    v.A[l] = on ? lv->A[l] : 1;
    v.HW[l] = on ? lv->HW[l] : 0;
    v.k[l] = on ? lv->k[l] : 0;
    v.first[l] = first;
    if (on) {
      CPM_CHECK_ARG(lv->A[l] >= 1 && lv->HW[l] >= 1 && (long)lv->A[l] * lv->HW[l] <= lv->row, "level %d: bad A / HW / row", l);
      CPM_CHECK_ARG(lv->k[l] >= 0 && lv->k[l] <= (long)lv->A[l] * lv->HW[l], "level %d: k out of range", l);
      int rc;
      if (need_obj && (rc = check_device_ptr(lv->d_objectness[l], "objectness")) != CPM_OK) return rc;
      if (need_reg && (rc = check_device_ptr(lv->d_regression[l], "regression")) != CPM_OK) return rc;
      if (need_reg && (rc = check_device_ptr(lv->d_anchors[l], "anchors")) != CPM_OK) return rc;
      if (need_reg) CPM_CHECK_ARG(((uintptr_t)lv->d_anchors[l] & 15) == 0, "anchors must be 16-byte aligned");
      first += (long)lv->num_images * lv->k[l];
    }
  }
  v.first[CPM_MAX_LEVELS] = first;
  for (int l = lv->num_levels; l < CPM_MAX_LEVELS; l++) v.first[l] = first;
  return CPM_OK;
}

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_rpn_flatten_objectness(const cpm_rpn_levels_t* levels, float* d_rows, void* stream) {
  RpnLevelsView v;
  int rc = make_levels_view(levels, v, true, false);
  if (rc != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_rows, "rows")) != CPM_OK) return rc;
  CPM_CHECK_ARG(v.num_levels * v.num_images < 65536, "too many (level, image) rows");
  dim3 grid((unsigned)((v.row + 1023) / 1024), (unsigned)(v.num_levels * v.num_images));
  rpn_flatten_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(v, d_rows);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

extern "C" int cpm_rpn_select_decode(const cpm_rpn_levels_t* levels, const int64_t* d_topk_idx, const float* d_topk_val,
                                     int64_t top_k, const float* d_image_wh, int64_t num_trash, const float* weights,
                                     float bbox_xform_clip, float min_size, float* d_boxes, float* d_scores,
                                     int32_t* d_segments_out, void* stream) {
  RpnLevelsView v;
  int rc = make_levels_view(levels, v, false, true);
  if (rc != CPM_OK) return rc;
  CPM_CHECK_ARG(num_trash >= 1 && top_k >= 1, "bad sizes");
  CPM_CHECK_ARG(weights != nullptr && weights[0] != 0.f && weights[1] != 0.f && weights[2] != 0.f && weights[3] != 0.f,
                "box coder weights must be non-zero");
  for (int l = 0; l < v.num_levels; l++) CPM_CHECK_ARG(v.k[l] <= top_k, "level %d: k exceeds the top-k row", l);
  const long M = v.first[CPM_MAX_LEVELS];
  if (M == 0) return CPM_OK;
  if ((rc = check_device_ptr(d_topk_idx, "topk_idx")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_topk_val, "topk_val")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_image_wh, "image_wh")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_boxes, "boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_scores, "scores")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_segments_out, "segments_out")) != CPM_OK) return rc;
  CPM_CHECK_ARG((((uintptr_t)d_boxes) & 15) == 0 && ((uintptr_t)d_image_wh & 7) == 0, "boxes / image_wh alignment");
  rpn_select_decode_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      v, (const long long*)d_topk_idx, d_topk_val, (int)top_k, (const float2*)d_image_wh, M, (int)num_trash, weights[0],
      weights[1], weights[2], weights[3], bbox_xform_clip, min_size, (float4*)d_boxes, d_scores, d_segments_out);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

extern "C" int cpm_rpn_decode(const float* d_deltas, const float* d_anchors, const int32_t* d_segments,
                              const float* d_segment_im_wh, int64_t M, int64_t num_segments, int64_t num_trash,
                              const float* weights, float bbox_xform_clip, float min_size, float* d_boxes,
                              int32_t* d_segments_out, void* stream) {
  CPM_CHECK_ARG(M >= 0 && num_segments >= 1 && num_trash >= 1 && num_segments + num_trash < (1LL << 30), "bad sizes");
  CPM_CHECK_ARG(weights != nullptr && weights[0] != 0.f && weights[1] != 0.f && weights[2] != 0.f && weights[3] != 0.f,
                "box coder weights must be non-zero");
  if (M == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_deltas, "deltas")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_anchors, "anchors")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_segments, "segments")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_segment_im_wh, "segment_im_wh")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_boxes, "boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_segments_out, "segments_out")) != CPM_OK) return rc;
  CPM_CHECK_ARG((((uintptr_t)d_deltas | (uintptr_t)d_anchors | (uintptr_t)d_boxes) & 15) == 0 &&
                    ((uintptr_t)d_segment_im_wh & 7) == 0,
                "deltas / anchors / boxes must be 16-byte aligned");
  rpn_decode_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)d_deltas, (const float4*)d_anchors, d_segments, (const float2*)d_segment_im_wh, M, (int)num_segments,
      (int)num_trash, weights[0], weights[1], weights[2], weights[3], bbox_xform_clip, min_size, (float4*)d_boxes, d_segments_out);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
