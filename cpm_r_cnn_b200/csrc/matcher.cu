// Box IoU matrix + Matcher on the device.
//
// Replaces, ahead of every subsample() of the heads,
//   boxlist_iou     pet/utils/data/structures/boxlist_ops.py:123-158   (N x M x 2 broadcast temporaries in torch)
//   Matcher         pet/rcnn/utils/matcher.py:52-112                   (max over dim 0, two masked writes, and for
//                   allow_low_quality_matches a max over dim 1 + `==` + nonzero() host sync + two index ops)
// cpm_box_iou writes the (N, M) matrix with one thread per element in the reference's operation order (areas with the
// +1 convention, inter / (area1 + area2 - inter)); cpm_matcher reads an (M, N) quality matrix twice: pass 1 takes every
// column's first maximum and every row's maximum (order-independent atomicMax on the non-negative float bits), pass 2
// applies the thresholds and restores the low-quality matches.  No host synchronisation; results are bit-identical to
// the torch ops (maxima are exact, the arithmetic is the same fp32 sequence without FMA contraction).
#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

__global__ void __launch_bounds__(256) box_iou_kernel(const float4* __restrict__ a, const float4* __restrict__ b, long N, long M,
                                                       float* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= N * M) return;
  const long n = i / M, m = i - n * M;
  const float4 p = a[n], q = b[m];
  const float area1 = (p.z - p.x + 1.0f) * (p.w - p.y + 1.0f);            // BoxList.area(), bounding_box.py:306-310
  const float area2 = (q.z - q.x + 1.0f) * (q.w - q.y + 1.0f);
  const float w = fmaxf((fminf(p.z, q.z) - fmaxf(p.x, q.x)) + 1.0f, 0.0f); // (rb - lt + TO_REMOVE).clamp(min=0)
  const float h = fmaxf((fminf(p.w, q.w) - fmaxf(p.y, q.y)) + 1.0f, 0.0f);
  const float inter = w * h;
  out[i] = inter / ((area1 + area2) - inter);
}

// pass 1: column maxima (first index wins, as torch.max(dim=0)) and row maxima
__global__ void __launch_bounds__(256) matcher_pass1(const float* __restrict__ q, long M, long N, float* __restrict__ col_val,
                                                      long long* __restrict__ col_idx, unsigned* __restrict__ row_max) {
  const long n = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float best = -INFINITY;
  long long bi = 0;
  for (long m = 0; m < M; m++) {
    const float v = n < N ? q[m * N + n] : 0.0f;
    if (n < N && v > best) {
      best = v;
      bi = m;
    }
    // qualities are >= 0 (IoU), so the unsigned order of the bits is the float order
    const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(n < N ? fmaxf(v, 0.0f) : 0.0f));
    if (lane == 0) atomicMax(row_max + m, wmax);
  }
  if (n < N) {
    col_val[n] = best;
    col_idx[n] = bi;
  }
}

// pass 2: thresholds (matcher.py:71-79) and set_low_quality_matches_ (:85-112)
__global__ void __launch_bounds__(256) matcher_pass2(const float* __restrict__ q, long M, long N, const float* __restrict__ col_val,
                                                      const unsigned* __restrict__ row_max, float high, float low, int allow_low,
                                                      long long* __restrict__ matches) {
  const long n = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float v = col_val[n];
  const long long all = matches[n];
  long long mt = all;
  if (v < low) mt = -1;                                    // BELOW_LOW_THRESHOLD
  else if (v < high) mt = -2;                              // BETWEEN_THRESHOLDS
  if (allow_low && mt < 0) {
    for (long m = 0; m < M; m++)
      if (__float_as_uint(q[m * N + n]) == row_max[m]) {   // this prediction ties the best overlap of ground truth m
        mt = all;
        break;
      }
  }
  matches[n] = mt;
}

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_box_iou(const float* d_boxes1, const float* d_boxes2, int64_t N, int64_t M, float* d_iou, void* stream) {
  CPM_CHECK_ARG(N >= 0 && M >= 0, "negative size");
  if (N == 0 || M == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_boxes1, "boxes1")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_boxes2, "boxes2")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_iou, "iou")) != CPM_OK) return rc;
  CPM_CHECK_ARG((((uintptr_t)d_boxes1 | (uintptr_t)d_boxes2) & 15) == 0, "boxes must be 16-byte aligned");
  const long total = (long)N * M;
  box_iou_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_boxes1,
                                                                                    (const float4*)d_boxes2, N, M, d_iou);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

extern "C" size_t cpm_matcher_workspace_bytes(int64_t M, int64_t N) {
  return ((size_t)(M > 0 ? M : 1) * sizeof(unsigned) + 255) / 256 * 256 + (size_t)(N > 0 ? N : 1) * sizeof(float);
}

extern "C" int cpm_matcher(const float* d_quality, int64_t M, int64_t N, float high_threshold, float low_threshold,
                           int allow_low_quality_matches, int64_t* d_matches, void* d_workspace, size_t workspace_bytes,
                           void* stream) {
  CPM_CHECK_ARG(M >= 1 && N >= 1, "empty match quality matrix (the reference raises ValueError, matcher.py:61-69)");
  CPM_CHECK_ARG(low_threshold <= high_threshold, "low_threshold > high_threshold");
  int rc;
  if ((rc = check_device_ptr(d_quality, "quality")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_matches, "matches")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_workspace, "workspace")) != CPM_OK) return rc;
  if (workspace_bytes < cpm_matcher_workspace_bytes(M, N)) {
    set_error("workspace too small");
    return CPM_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* row_max = (unsigned*)d_workspace;
  float* col_val = (float*)((char*)d_workspace + ((size_t)M * sizeof(unsigned) + 255) / 256 * 256);
  CPM_CHECK_CUDA(cudaMemsetAsync(row_max, 0, (size_t)M * sizeof(unsigned), st));
  const unsigned blocks = (unsigned)((N + 255) / 256);
  matcher_pass1<<<blocks, 256, 0, st>>>(d_quality, M, N, col_val, (long long*)d_matches, row_max);
  CPM_CHECK_LAUNCH();
  matcher_pass2<<<blocks, 256, 0, st>>>(d_quality, M, N, col_val, row_max, high_threshold, low_threshold,
                                        allow_low_quality_matches, (long long*)d_matches);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
