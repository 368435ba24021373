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

}  // namespace cpm

using namespace cpm;

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
