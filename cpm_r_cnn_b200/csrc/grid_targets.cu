// Grid-point training targets rasterised on the device.
//
// Replaces GridLossComputation.prepare_target (pet/rcnn/modeling/grid_cascade_rcnn/loss.py:178-258): a Python triple loop
// over (RoI, grid point, disk pixel) on CPU tensors that fills a (R, P, 56, 56) map, crops every point's 28x28 sub-region
// (Grid R-CNN Plus) and uploads the result.  Here one CTA writes one (RoI, point) sub-map directly: the box extension,
// the grid point, its integer map position (truncation of the fp32 expression, as int() does on the 0-dim tensors of the
// reference) and the disk test are evaluated per CTA / per pixel; nothing touches the host.  fp32 operation order follows
// the torch expressions (no FMA contraction in this library).
#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

struct GridSub {
  int v[2 * 64];
};

__global__ void __launch_bounds__(256) grid_targets_kernel(const float4* __restrict__ pos, const float4* __restrict__ gt, int P,
                                                            int gs, int map_size, int half, float ratio, int radius,
                                                            int refine, GridSub sub, float* __restrict__ out) {
  const long r = blockIdx.x / P;
  const int j = blockIdx.x % P;
  const float4 b = pos[r], g = gt[r];
  // loss.py:187-193: the RoI extended by mapping_ratio on every side
  const float x1 = b.x - ratio * ((b.z - b.x) / 2.0f), y1 = b.y - ratio * ((b.w - b.y) / 2.0f);
  const float x2 = b.z + ratio * ((b.z - b.x) / 2.0f), y2 = b.w + ratio * ((b.w - b.y) / 2.0f);
  const float bw = x2 - x1, bh = y2 - y1;
  float* o = out + ((size_t)r * P + j) * half * half;
  bool skip = bw <= (float)gs || bh <= (float)gs;          // :216-218 "ignore small bboxes"
  int cx = 0, cy = 0;
  if (!skip) {
    const int x_idx = j / gs, y_idx = j % gs;              // :206-209 interpolation factors (python floats -> fp32 scalars)
    const float fx = (float)(1.0 - (double)x_idx / (double)(gs - 1)), fy = (float)(1.0 - (double)y_idx / (double)(gs - 1));
    const float omfx = (float)(1.0 - (1.0 - (double)x_idx / (double)(gs - 1))), omfy = (float)(1.0 - (1.0 - (double)y_idx / (double)(gs - 1)));
    const float gx = fx * g.x + omfx * g.z, gy = fy * g.y + omfy * g.w;         // :222-225
    cx = (int)((gx - x1) / bw * (float)map_size);          // :227-230, int() truncates toward zero
    cy = (int)((gy - y1) / bh * (float)map_size);
  }
  const int sx = sub.v[2 * j], sy = sub.v[2 * j + 1];
  const int r2 = radius * radius;
  int rx = -1, ry = -1;                                    // TARGET_REFINE (:238-250): a centre outside the map marks its
  if (!skip && refine && (cx < 0 || cx >= map_size || cy < 0 || cy >= map_size)) {     // clamped pixel
    rx = min(max(cx, 0), map_size - 1);
    ry = min(max(cy, 0), map_size - 1);
  }
  for (int e = threadIdx.x; e < half * half; e += blockDim.x) {
    const int x = sx + e % half, y = sy + e / half;        // position in the whole 4x map
    float v = 0.f;
    if (!skip) {
      const int dx = x - cx, dy = y - cy;
      if (dx >= -radius && dx <= radius && dy >= -radius && dy <= radius && dx * dx + dy * dy <= r2) v = 1.f;
      if (x == rx && y == ry) v = 1.f;
    }
    o[e] = v;
  }
}

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_grid_targets(const float* d_pos_boxes, const float* d_gt_boxes, int64_t R, int grid_points, int map_size,
                                const int32_t* sub_xy, float mapping_ratio, int pos_radius, int target_refine,
                                float* d_targets, void* stream) {
  CPM_CHECK_ARG(R >= 0 && R * (int64_t)grid_points < (1LL << 31), "R out of range");
  CPM_CHECK_ARG(grid_points >= 4 && grid_points <= 64, "grid_points must be in [4, 64]");
  int gs = 1;
  while (gs * gs < grid_points) gs++;
  CPM_CHECK_ARG(gs * gs == grid_points, "grid_points must be a square number");
  CPM_CHECK_ARG(map_size >= 4 && map_size % 4 == 0 && pos_radius >= 0 && sub_xy != nullptr, "bad map_size / radius / sub_xy");
  if (R == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_pos_boxes, "pos_boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_gt_boxes, "gt_boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_targets, "targets")) != CPM_OK) return rc;
  CPM_CHECK_ARG((((uintptr_t)d_pos_boxes | (uintptr_t)d_gt_boxes) & 15) == 0, "boxes must be 16-byte aligned");
  GridSub sub;
  for (int i = 0; i < 2 * grid_points; i++) sub.v[i] = sub_xy[i];
  for (int i = 2 * grid_points; i < 128; i++) sub.v[i] = 0;
  const int half = map_size / 4 * 2;
  grid_targets_kernel<<<(unsigned)(R * grid_points), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)d_pos_boxes, (const float4*)d_gt_boxes, grid_points, gs, map_size, half, mapping_ratio, pos_radius,
      target_refine, sub, d_targets);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
