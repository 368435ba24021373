// Shared helpers for the cpm_ops kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cpm_ops.h"

namespace cpm {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CPM_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      cpm::set_error(__VA_ARGS__);            \
      return CPM_ERR_INVALID_ARG;             \
    }                                         \
  } while (0)

#define CPM_CHECK_CUDA(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      cpm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CPM_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// after a <<<>>> launch
#define CPM_CHECK_LAUNCH()                   \
  do {                                       \
    cpm::count_launch();                     \
    CPM_CHECK_CUDA(cudaGetLastError());      \
  } while (0)

// Kernel-side view of cpm_pyramid_t (passed by value as a kernel parameter).
struct PyramidView {
  int num_levels, batch, channels, dtype, layout;
  void* ptr[CPM_MAX_LEVELS];
  int H[CPM_MAX_LEVELS];
  int W[CPM_MAX_LEVELS];
  float scale[CPM_MAX_LEVELS];
};

struct MapperView {
  float k_min, k_max, inv_s0, lvl0, eps;
};

inline PyramidView make_view(const cpm_pyramid_t* p) {
  PyramidView v;
  v.num_levels = p->num_levels;
  v.batch = p->batch;
  v.channels = p->channels;
  v.dtype = p->dtype;
  v.layout = p->layout;
  for (int i = 0; i < CPM_MAX_LEVELS; i++) {
    v.ptr[i] = i < p->num_levels ? p->d_level[i] : nullptr;
    v.H[i] = i < p->num_levels ? p->height[i] : 0;
    v.W[i] = i < p->num_levels ? p->width[i] : 0;
    v.scale[i] = i < p->num_levels ? p->spatial_scale[i] : 0.f;
  }
  return v;
}

inline MapperView make_view(const cpm_level_mapper_t* m) {
  MapperView v{0.f, 0.f, 1.f / 224.f, 4.f, 1e-6f};
  if (m) {
    v.k_min = m->k_min;
    v.k_max = m->k_max;
    v.inv_s0 = 1.0f / m->canonical_scale;
    v.lvl0 = m->canonical_level;
    v.eps = m->eps;
  }
  return v;
}

// LevelMapper.__call__ (pet/rcnn/utils/poolers.py:29-40) for one box; torch-on-CUDA op order, no contraction.
__device__ __forceinline__ int fpn_level(float x1, float y1, float x2, float y2, const MapperView& m) {
  float area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.0f), __fadd_rn(__fsub_rn(y2, y1), 1.0f));
  float s = sqrtf(area);
  float t = floorf(__fadd_rn(m.lvl0, log2f(__fadd_rn(__fmul_rn(s, m.inv_s0), m.eps))));
  t = fminf(fmaxf(t, m.k_min), m.k_max);   // NaN -> k_min (torch.clamp would propagate NaN; not reachable for finite boxes)
  return (int)t - (int)m.k_min;
}

// Geometry of one RoI on its level: ROIAlign_cuda.cu:199-230.
template <typename T>
struct RoiGeo {
  T start_w, start_h, bin_w, bin_h;
  int gh, gw;       // sampling grid per bin
  int b;            // batch index
};

template <typename T>
__device__ __forceinline__ RoiGeo<T> roi_geometry(const T* roi, T scale, int PH, int PW, int sr, bool aligned) {
  RoiGeo<T> g;
  g.b = (int)roi[0];
  T off = aligned ? (T)0.5 : (T)0.0;
  g.start_w = roi[1] * scale - off;
  g.start_h = roi[2] * scale - off;
  T end_w = roi[3] * scale - off;
  T end_h = roi[4] * scale - off;
  T rw = end_w - g.start_w;
  T rh = end_h - g.start_h;
  if (!aligned) {
    rw = max(rw, (T)1.);
    rh = max(rh, (T)1.);
  }
  g.bin_h = rh / (T)PH;
  g.bin_w = rw / (T)PW;
  g.gh = sr > 0 ? sr : (int)ceil(rh / (T)PH);
  g.gw = sr > 0 ? sr : (int)ceil(rw / (T)PW);
  return g;
}

// One sample coordinate along one axis resolved into (low index, high index, low weight, high weight).
// Restates the per-axis half of bilinear_interpolate (ROIAlign_cuda.cu:36-86): the validity test, the clamp to
// [0, size-1] and the weights all factor per axis (w1 = hy*hx, ...), which is what the separable kernels use.
struct AxisTap {
  int lo, hi;
  float wlo, whi;   // (1 - l), l ; both 0 when the coordinate is out of range
  int valid;
};

__device__ __forceinline__ AxisTap axis_tap(float v, int size) {
  AxisTap t;
  if (v < -1.0f || v > (float)size) {
    t.lo = t.hi = 0;
    t.wlo = t.whi = 0.f;
    t.valid = 0;
    return t;
  }
  if (v <= 0.f) v = 0.f;
  int lo = (int)v;
  int hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  float l = v - (float)lo;
  t.lo = lo;
  t.hi = hi;
  t.whi = l;
  t.wlo = 1.f - l;
  t.valid = 1;
  return t;
}

__device__ __forceinline__ int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace cpm
