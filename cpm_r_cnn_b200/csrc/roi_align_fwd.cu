// RoIAlign forward for sm_100a.
//
// Replaces ROIAlign_forward_cuda (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:367-425, kernel :178-256) and, for a
// multi-level pyramid, the per-level nonzero/gather/index_put loop of Pooler.forward (pet/rcnn/utils/poolers.py:117-131).
//
// Kernels, in dispatch order:
//   * roi_align_fwd_cols (roi_align_fwd_cols.cu): the hot path -- the CPM poolers (7x7 / 14x14, sampling_ratio 1 or 2) on
//     an NHWC fp32 or bf16 pyramid, both pooled layouts.
//   * roi_align_fwd_nhwc_any: any pooled size / fixed or adaptive sampling grid on an NHWC fp32 pyramid (C % 4 == 0).  One
//     CTA per (RoI, 128-channel chunk), a lane owns 4 consecutive channels, every bilinear tap is one coalesced 512-byte
//     row; the pooled block is transposed through shared memory into the (K, C, PH, PW) output.
//   * roi_align_fwd_nhwc_bf16: the same for a bf16 pyramid (lane = 8 channels).
//   * roi_align_fwd_generic<T>: one thread per output element, any layout / dtype / interpolation / sampling grid; the
//     arithmetic is the reference kernel's, statement by statement (fp64, nearest, NCHW, odd channel counts).
//   The FPN level of each RoI (LevelMapper, poolers.py:29-40) is evaluated inside every kernel.
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cpm {

// ------------------------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T bilinear_ref(const T* f, long sy, long sx, int H, int W, T y, T x) {
  // ROIAlign_cuda.cu:36-86
  if (y < (T)-1.0 || y > (T)H || x < (T)-1.0 || x > (T)W) return (T)0;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) {
    y_high = y_low = H - 1;
    y = (T)y_low;
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= W - 1) {
    x_high = x_low = W - 1;
    x = (T)x_low;
  } else {
    x_high = x_low + 1;
  }
  T ly = y - y_low, lx = x - x_low, hy = (T)1. - ly, hx = (T)1. - lx;
  T v1 = f[y_low * sy + x_low * sx], v2 = f[y_low * sy + x_high * sx];
  T v3 = f[y_high * sy + x_low * sx], v4 = f[y_high * sy + x_high * sx];
  T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

template <typename T>
__device__ __forceinline__ T nearest_ref(const T* f, long sy, long sx, int H, int W, T y, T x) {
  // ROIAlign_cuda.cu:14-33
  if (y < (T)-0.5 || y >= (T)H - (T)0.5 || x < (T)-0.5 || x >= (T)W - (T)0.5) return (T)0;
  int x_low = (int)round(x), y_low = (int)round(y);
  return f[y_low * sy + x_low * sx];
}

template <typename T>
__device__ __forceinline__ int roi_level(const T* roi, const PyramidView& pv, const MapperView& mp, const int* lv,
                                         long n) {
  if (pv.num_levels == 1) return 0;
  if (lv) return lv[n];
  return fpn_level((float)roi[1], (float)roi[2], (float)roi[3], (float)roi[4], mp);
}

template <typename T>
__global__ void __launch_bounds__(256) roi_align_fwd_generic(PyramidView pv, const T* __restrict__ rois, long K, int PH,
                                                              int PW, int sr, int aligned, int interp, MapperView mp,
                                                              const int* __restrict__ roi_levels, T* __restrict__ out) {
  const int C = pv.channels;
  const long total = K * C * PH * PW;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int pw = idx % PW;
    int ph = (idx / PW) % PH;
    int c = (idx / PW / PH) % C;
    long n = idx / PW / PH / C;
    const T* roi = rois + 5 * n;
    int l = roi_level(roi, pv, mp, roi_levels, n);
    if (l < 0 || l >= pv.num_levels) {   // caller-provided level out of range: defined as zeros
      out[idx] = (T)0;
      continue;
    }
    const int H = pv.H[l], W = pv.W[l];
    RoiGeo<T> g = roi_geometry<T>(roi, (T)pv.scale[l], PH, PW, sr, aligned != 0);
    if (g.b < 0 || g.b >= pv.batch) {    // image index outside the batch: defined as zeros (every kernel of the library)
      out[idx] = (T)0;
      continue;
    }
    long sC, sY, sX;
    if (pv.layout == CPM_LAYOUT_NCHW) {
      sC = (long)H * W; sY = W; sX = 1;
    } else {
      sC = 1; sY = (long)W * C; sX = C;
    }
    const T* f = (const T*)pv.ptr[l] + (long)g.b * C * H * W + c * sC;
    int cnt_i = g.gh * g.gw;
    const T count = (T)(cnt_i > 1 ? cnt_i : 1);
    T acc = 0;
    for (int iy = 0; iy < g.gh; iy++) {
      const T y = g.start_h + ph * g.bin_h + static_cast<T>(iy + .5f) * g.bin_h / static_cast<T>(g.gh);
      for (int ix = 0; ix < g.gw; ix++) {
        const T x = g.start_w + pw * g.bin_w + static_cast<T>(ix + .5f) * g.bin_w / static_cast<T>(g.gw);
        acc += interp == CPM_INTERP_BILINEAR ? bilinear_ref<T>(f, sY, sX, H, W, y, x)
                                             : nearest_ref<T>(f, sY, sX, H, W, y, x);
      }
    }
    out[idx] = acc / count;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NHWC hot path
// ------------------------------------------------------------------------------------------------------------------
constexpr int kChunk = 128;          // channels per CTA: 32 lanes x float4

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// 4 channels of one bilinear sample: w1*v1 + w2*v2 + w3*v3 + w4*v4 in the reference's left-to-right order,
// contracted the way nvcc contracts ROIAlign_cuda.cu:84 (mul, fma, fma, fma).
__device__ __forceinline__ float tap4(float w1, float v1, float w2, float v2, float w3, float v3, float w4, float v4) {
  return fmaf(w4, v4, fmaf(w3, v3, fmaf(w2, v2, w1 * v1)));
}

// Shared-memory staging tile: channel-major rows of `S` floats (S = PH*PW rounded up to odd) with a one-word skew
// every 32 channels so that the 4-channel-per-lane scatter is bank-conflict free.
__device__ __forceinline__ int stage_off(int c, int S) { return c * S + (c >> 5); }

// Runtime-shaped kernel: any PH, PW, sampling grid (also adaptive).  One accumulator (4 channels) per lane at a time.
__global__ void __launch_bounds__(256) roi_align_fwd_nhwc_any(PyramidView pv, const float* __restrict__ rois, int PH, int PW,
                                                               int sr, int aligned, MapperView mp,
                                                               const int* __restrict__ roi_levels, float* __restrict__ out,
                                                               int chunks, int S) {
  extern __shared__ float tile[];
  const int C = pv.channels;
  const long n = blockIdx.x / chunks;
  const int c0 = (blockIdx.x % chunks) * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cc = min(kChunk, C - c0);          // channels in this chunk (multiple of 4)
  const float* roi = rois + 5 * n;
  const int l = roi_level(roi, pv, mp, roi_levels, n);
  const int PP = PH * PW;
  const int bidx = (int)roi[0];
  if (l >= 0 && l < pv.num_levels && bidx >= 0 && bidx < pv.batch) {
    const int H = pv.H[l], W = pv.W[l];
    RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, sr, aligned != 0);
    const int cnt_i = g.gh * g.gw;
    const float count = (float)(cnt_i > 1 ? cnt_i : 1);
    const bool active = 4 * lane < cc;
    const float4* f = (const float4*)((const float*)pv.ptr[l] + (long)g.b * H * W * C + c0) + lane;
    const long C4 = C >> 2;
    for (int bin = warp; bin < PP; bin += nwarps) {
      const int ph = bin / PW, pw = bin % PW;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int iy = 0; iy < g.gh; iy++) {
        const float y = g.start_h + ph * g.bin_h + static_cast<float>(iy + .5f) * g.bin_h / static_cast<float>(g.gh);
        const AxisTap ty = axis_tap(y, H);
        for (int ix = 0; ix < g.gw; ix++) {
          const float x = g.start_w + pw * g.bin_w + static_cast<float>(ix + .5f) * g.bin_w / static_cast<float>(g.gw);
          const AxisTap tx = axis_tap(x, W);
          if (!(ty.valid && tx.valid) || !active) continue;
          const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
          const float4 v1 = ldg_f4(f + ((long)ty.lo * W + tx.lo) * C4);
          const float4 v2 = ldg_f4(f + ((long)ty.lo * W + tx.hi) * C4);
          const float4 v3 = ldg_f4(f + ((long)ty.hi * W + tx.lo) * C4);
          const float4 v4 = ldg_f4(f + ((long)ty.hi * W + tx.hi) * C4);
          acc.x += tap4(w1, v1.x, w2, v2.x, w3, v3.x, w4, v4.x);
          acc.y += tap4(w1, v1.y, w2, v2.y, w3, v3.y, w4, v4.y);
          acc.z += tap4(w1, v1.z, w2, v2.z, w3, v3.z, w4, v4.z);
          acc.w += tap4(w1, v1.w, w2, v2.w, w3, v3.w, w4, v4.w);
        }
      }
      if (active) {
        const int c = 4 * lane;
        tile[stage_off(c + 0, S) + bin] = acc.x / count;
        tile[stage_off(c + 1, S) + bin] = acc.y / count;
        tile[stage_off(c + 2, S) + bin] = acc.z / count;
        tile[stage_off(c + 3, S) + bin] = acc.w / count;
      }
    }
  } else {
    for (int e = threadIdx.x; e < cc * PP; e += blockDim.x) tile[stage_off(e / PP, S) + e % PP] = 0.f;
  }
  __syncthreads();
  // the chunk's pooled block out[n, c0:c0+cc, :, :] is contiguous: stream it out coalesced
  float* o = out + ((long)n * C + c0) * PP;
  const int total = cc * PP;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e / PP, b = e - c * PP;
    o[e] = tile[stage_off(c, S) + b];
  }
}

// bf16 storage, fp32 arithmetic (new on this path; the reference computes half inputs in fp32 through
// apex.amp.float_function, roi_align.py:76, after a full-size cast of the feature maps).  Same structure as the kernel
// above with a lane owning 8 channels: every tap is one 16-byte bf16x8 load, a warp covers 256 channels; the pooled block
// is rounded to bf16 once, at the end (round-to-nearest-even), and leaves through a bf16 shared-memory tile.
constexpr int kChunkB = 256;

__global__ void __launch_bounds__(256) roi_align_fwd_nhwc_bf16(PyramidView pv, const float* __restrict__ rois, int PH, int PW,
                                                                int sr, int aligned, MapperView mp,
                                                                const int* __restrict__ roi_levels,
                                                                __nv_bfloat16* __restrict__ out, int chunks, int S) {
  extern __shared__ __align__(16) unsigned char tile_raw[];
  __nv_bfloat16* tileb = reinterpret_cast<__nv_bfloat16*>(tile_raw);
  const int C = pv.channels;
  const long n = blockIdx.x / chunks;
  const int c0 = (blockIdx.x % chunks) * kChunkB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cc = min(kChunkB, C - c0);         // channels in this chunk (multiple of 8)
  const float* roi = rois + 5 * n;
  const int l = roi_level(roi, pv, mp, roi_levels, n);
  const int PP = PH * PW;
  bool done = false;
  if (l >= 0 && l < pv.num_levels) {
    const int H = pv.H[l], W = pv.W[l];
    RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, sr, aligned != 0);
    if (g.b >= 0 && g.b < pv.batch) {
      done = true;
      const int cnt_i = g.gh * g.gw;
      const float count = (float)(cnt_i > 1 ? cnt_i : 1);
      const bool active = 8 * lane < cc;
      const uint4* f = reinterpret_cast<const uint4*>((const __nv_bfloat16*)pv.ptr[l] + (long)g.b * H * W * C + c0) + lane;
      const long C8 = C >> 3;
      for (int bin = warp; bin < PP; bin += nwarps) {
        const int ph = bin / PW, pw = bin % PW;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
        for (int iy = 0; iy < g.gh; iy++) {
          const float y = g.start_h + ph * g.bin_h + static_cast<float>(iy + .5f) * g.bin_h / static_cast<float>(g.gh);
          const AxisTap ty = axis_tap(y, H);
          for (int ix = 0; ix < g.gw; ix++) {
            const float x = g.start_w + pw * g.bin_w + static_cast<float>(ix + .5f) * g.bin_w / static_cast<float>(g.gw);
            const AxisTap tx = axis_tap(x, W);
            if (!(ty.valid && tx.valid) || !active) continue;
            const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
            const uint4 v1 = __ldg(f + ((long)ty.lo * W + tx.lo) * C8);
            const uint4 v2 = __ldg(f + ((long)ty.lo * W + tx.hi) * C8);
            const uint4 v3 = __ldg(f + ((long)ty.hi * W + tx.lo) * C8);
            const uint4 v4 = __ldg(f + ((long)ty.hi * W + tx.hi) * C8);
            const unsigned a1[4] = {v1.x, v1.y, v1.z, v1.w}, a2[4] = {v2.x, v2.y, v2.z, v2.w};
            const unsigned a3[4] = {v3.x, v3.y, v3.z, v3.w}, a4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {      // a 32-bit word holds two bf16: the low half is the even channel
              acc[2 * k] += tap4(w1, __uint_as_float(a1[k] << 16), w2, __uint_as_float(a2[k] << 16),
                                 w3, __uint_as_float(a3[k] << 16), w4, __uint_as_float(a4[k] << 16));
              acc[2 * k + 1] += tap4(w1, __uint_as_float(a1[k] & 0xffff0000u), w2, __uint_as_float(a2[k] & 0xffff0000u),
                                     w3, __uint_as_float(a3[k] & 0xffff0000u), w4, __uint_as_float(a4[k] & 0xffff0000u));
            }
          }
        }
        if (active) {
#pragma unroll
          for (int k = 0; k < 8; k++) tileb[(8 * lane + k) * S + bin] = __float2bfloat16_rn(acc[k] / count);
        }
      }
    }
  }
  if (!done)
    for (int e = threadIdx.x; e < cc * PP; e += blockDim.x) tileb[(e / PP) * S + e % PP] = __float2bfloat16_rn(0.f);
  __syncthreads();
  __nv_bfloat16* o = out + ((long)n * C + c0) * PP;
  const int total = cc * PP;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e / PP, b = e - c * PP;
    o[e] = tileb[c * S + b];
  }
}

int check_device_ptr(const void* p, const char* what);
bool fwd_cols_supported(const cpm_pyramid_t* feat, int pooled_h, int pooled_w, int sampling_ratio, const void* d_out);
int launch_fwd_cols(const PyramidView& pv, const float* rois, long K, int P, int G, int aligned, const MapperView& mp,
                    const int* lv, void* out, int out_channels_last, cudaStream_t st);

bool fwd_rows_supported(const cpm_pyramid_t* feat, int pooled_h, int pooled_w, int sampling_ratio, const void* d_out);
int launch_fwd_rows(const cpm_pyramid_t* feat, const PyramidView& pv, const float* rois, long K, int P, int G, int aligned,
                    const MapperView& mp, const int* lv, float* out, cudaStream_t st);

// CPM_FWD_IMPL=rows|cols forces one of the two kernels of the CPM poolers for calls that leave the choice to the library
// (default: the row-streaming kernel where it applies -- fp32, (K,C,PH,PW) output -- else the column-table kernel)
static int fwd_impl_env() {
  static const int v = [] {
    const char* e = getenv("CPM_FWD_IMPL");
    if (e == nullptr) return 0;
    if (strcmp(e, "rows") == 0) return CPM_FWD_ROWS;
    if (strcmp(e, "cols") == 0) return CPM_FWD_COLS;
    return 0;
  }();
  return v;
}

int check_pyramid(const cpm_pyramid_t* p, const char* what) {
  CPM_CHECK_ARG(p != nullptr, "%s is NULL", what);
  CPM_CHECK_ARG(p->num_levels >= 1 && p->num_levels <= CPM_MAX_LEVELS, "%s: num_levels %d not in [1,%d]", what,
                p->num_levels, CPM_MAX_LEVELS);
  CPM_CHECK_ARG(p->batch >= 0 && p->channels >= 1, "%s: bad batch/channels", what);
  CPM_CHECK_ARG(p->dtype == CPM_F32 || p->dtype == CPM_F64 || p->dtype == CPM_BF16, "%s: unknown dtype %d", what,
                p->dtype);
  CPM_CHECK_ARG(p->layout == CPM_LAYOUT_NCHW || p->layout == CPM_LAYOUT_NHWC, "%s: unknown layout %d", what, p->layout);
  for (int l = 0; l < p->num_levels; l++) {
    CPM_CHECK_ARG(p->height[l] >= 1 && p->width[l] >= 1, "%s: level %d has empty spatial size", what, l);
    if (p->batch > 0) {
      char nm[64];
      snprintf(nm, sizeof(nm), "%s level %d", what, l);
      int rc = check_device_ptr(p->d_level[l], nm);
      if (rc != CPM_OK) return rc;
    }
  }
  return CPM_OK;
}

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_roi_align_forward(const cpm_pyramid_t* feat, const void* d_rois, int64_t K, int pooled_h, int pooled_w,
                                     int sampling_ratio, int aligned, int interpolation, const cpm_level_mapper_t* mapper,
                                     const int32_t* d_roi_levels, int impl, void* d_out, void* stream) {
  return cpm_roi_align_forward_ex(feat, d_rois, K, pooled_h, pooled_w, sampling_ratio, aligned, interpolation, mapper,
                                  d_roi_levels, impl, CPM_POOLED_KCHW, d_out, stream);
}

extern "C" int cpm_roi_align_forward_ex(const cpm_pyramid_t* feat, const void* d_rois, int64_t K, int pooled_h, int pooled_w,
                                        int sampling_ratio, int aligned, int interpolation, const cpm_level_mapper_t* mapper,
                                        const int32_t* d_roi_levels, int impl, int pooled_layout, void* d_out, void* stream) {
  int rc = check_pyramid(feat, "feat");
  if (rc != CPM_OK) return rc;
  CPM_CHECK_ARG(pooled_layout == CPM_POOLED_KCHW || pooled_layout == CPM_POOLED_KHWC, "unknown pooled layout %d", pooled_layout);
  CPM_CHECK_ARG(K >= 0, "K < 0");
  CPM_CHECK_ARG(pooled_h >= 1 && pooled_w >= 1, "pooled size must be positive");
  CPM_CHECK_ARG(interpolation == CPM_INTERP_BILINEAR || interpolation == CPM_INTERP_NEAREST,
                "unknown interpolation method %d", interpolation);
  CPM_CHECK_ARG(feat->num_levels == 1 || mapper != nullptr || d_roi_levels != nullptr,
                "a multi-level pyramid needs a level mapper or per-RoI levels");
  if (K == 0) return CPM_OK;   // ROIAlign_cuda.cu:401-404
  if ((rc = check_device_ptr(d_rois, "rois")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_out, "out")) != CPM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PyramidView pv = make_view(feat);
  MapperView mp = make_view(mapper);
  const long total = (long)K * feat->channels * pooled_h * pooled_w;

  bool nhwc_ok = feat->layout == CPM_LAYOUT_NHWC && feat->dtype == CPM_F32 && interpolation == CPM_INTERP_BILINEAR &&
                 feat->channels % 4 == 0 && (long)K * ((feat->channels + kChunk - 1) / kChunk) < (1L << 31);
  for (int l = 0; nhwc_ok && l < feat->num_levels; l++) nhwc_ok = ((uintptr_t)feat->d_level[l] & 15) == 0;
  nhwc_ok = nhwc_ok && ((uintptr_t)d_out & 15) == 0;
  const int PP = pooled_h * pooled_w;
  const size_t smem_any = (size_t)(kChunk * (PP | 1) + 8) * sizeof(float);
  if (smem_any > 200 * 1024) nhwc_ok = false;
  // a bf16 pyramid (bf16 pooled output) takes the same column-table kernel with 8-byte taps
  const bool bf16_ok = feat->layout == CPM_LAYOUT_NHWC && feat->dtype == CPM_BF16 && interpolation == CPM_INTERP_BILINEAR &&
                       (long)K * ((feat->channels + 63) / 64) < (1L << 31);
  const bool cols_ok = (nhwc_ok || bf16_ok) && interpolation == CPM_INTERP_BILINEAR &&
                       fwd_cols_supported(feat, pooled_h, pooled_w, sampling_ratio, d_out);
  const bool rows_ok = nhwc_ok && pooled_layout == CPM_POOLED_KCHW &&
                       fwd_rows_supported(feat, pooled_h, pooled_w, sampling_ratio, d_out);
  if (impl == CPM_FWD_ROWS && !rows_ok) {
    set_error("CPM_FWD_ROWS needs an NHWC fp32 pyramid, a (K,C,PH,PW) output, bilinear interpolation, a 7x7 or 14x14 pooler, "
              "sampling_ratio 1 or 2, C %% 128 == 0 (7x7) / C %% 64 == 0 (14x14) and a driver with cuTensorMapEncodeTiled");
    return CPM_ERR_UNSUPPORTED;
  }
  if (rows_ok && (impl == CPM_FWD_ROWS || (impl == CPM_FWD_AUTO && fwd_impl_env() != CPM_FWD_COLS))) {
    rc = launch_fwd_rows(feat, pv, (const float*)d_rois, K, pooled_h, sampling_ratio, aligned, mp, d_roi_levels, (float*)d_out, st);
    if (rc != CPM_ERR_UNSUPPORTED || impl == CPM_FWD_ROWS) return rc;
  }
  if (impl == CPM_FWD_COLS && !cols_ok) {
    set_error("CPM_FWD_COLS needs an NHWC fp32 pyramid, bilinear interpolation, a 7x7 or 14x14 pooler, sampling_ratio 1 or 2 "
              "and C %% 128 == 0 (7x7) / C %% 64 == 0 (14x14)");
    return CPM_ERR_UNSUPPORTED;
  }
  if (cols_ok && (impl == CPM_FWD_AUTO || impl == CPM_FWD_COLS))
    return launch_fwd_cols(pv, (const float*)d_rois, K, pooled_h, sampling_ratio, aligned, mp, d_roi_levels, d_out,
                           pooled_layout == CPM_POOLED_KHWC, st);
  if (pooled_layout == CPM_POOLED_KHWC) {
    set_error("a channels-last pooled output (CPM_POOLED_KHWC) is produced by the column-table kernel only: NHWC fp32 pyramid, "
              "bilinear, 7x7 or 14x14 pooler, sampling_ratio 1 or 2, C %% 128 == 0 (7x7) / C %% 64 == 0 (14x14)");
    return CPM_ERR_UNSUPPORTED;
  }
  if (impl == CPM_FWD_NHWC && !nhwc_ok) {
    set_error("CPM_FWD_NHWC needs an NHWC fp32 pyramid, bilinear interpolation, C %% 4 == 0 and 16-byte aligned maps");
    return CPM_ERR_UNSUPPORTED;
  }
  if (nhwc_ok && impl != CPM_FWD_GENERIC) {
    const int chunks = (feat->channels + kChunk - 1) / kChunk;
    static thread_local int configured_dev = -1;
    int dev;
    CPM_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      CPM_CHECK_CUDA(
          cudaFuncSetAttribute(roi_align_fwd_nhwc_any, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured_dev = dev;
    }
    roi_align_fwd_nhwc_any<<<(unsigned)(K * chunks), 256, smem_any, st>>>(pv, (const float*)d_rois, pooled_h, pooled_w,
                                                                          sampling_ratio, aligned, mp, d_roi_levels,
                                                                          (float*)d_out, chunks, PP | 1);
    CPM_CHECK_LAUNCH();
    return CPM_OK;
  }
  if (feat->dtype == CPM_BF16) {
    bool ok = feat->layout == CPM_LAYOUT_NHWC && interpolation == CPM_INTERP_BILINEAR && feat->channels % 8 == 0 &&
              pooled_layout == CPM_POOLED_KCHW && ((uintptr_t)d_out & 1) == 0;
    for (int l = 0; ok && l < feat->num_levels; l++) ok = ((uintptr_t)feat->d_level[l] & 15) == 0;
    const int S = PP | 1;
    const int chunksb = (feat->channels + kChunkB - 1) / kChunkB;
    const size_t smem_b = (size_t)kChunkB * S * sizeof(__nv_bfloat16);
    if (!ok || smem_b > 200 * 1024 || (long)K * chunksb >= (1L << 31)) {
      set_error("bf16 pyramids need layout NHWC, bilinear interpolation, C %% 8 == 0, 16-byte aligned maps, a (K,C,PH,PW) "
                "bf16 output and PH*PW <= 399 (rois stay fp32)");
      return CPM_ERR_UNSUPPORTED;
    }
    static thread_local int configured_dev_b = -1;
    int devb;
    CPM_CHECK_CUDA(cudaGetDevice(&devb));
    if (configured_dev_b != devb) {
      CPM_CHECK_CUDA(cudaFuncSetAttribute(roi_align_fwd_nhwc_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured_dev_b = devb;
    }
    roi_align_fwd_nhwc_bf16<<<(unsigned)(K * chunksb), 256, smem_b, st>>>(pv, (const float*)d_rois, pooled_h, pooled_w,
                                                                           sampling_ratio, aligned, mp, d_roi_levels,
                                                                           (__nv_bfloat16*)d_out, chunksb, S);
    CPM_CHECK_LAUNCH();
    return CPM_OK;
  }
  const int threads = 256;
  long blocks = (total + threads - 1) / threads;
  if (blocks > 148L * 64) blocks = 148L * 64;
  if (feat->dtype == CPM_F32)
    roi_align_fwd_generic<float><<<(unsigned)blocks, threads, 0, st>>>(pv, (const float*)d_rois, K, pooled_h, pooled_w,
                                                                       sampling_ratio, aligned, interpolation, mp,
                                                                       d_roi_levels, (float*)d_out);
  else
    roi_align_fwd_generic<double><<<(unsigned)blocks, threads, 0, st>>>(pv, (const double*)d_rois, K, pooled_h,
                                                                        pooled_w, sampling_ratio, aligned, interpolation,
                                                                        mp, d_roi_levels, (double*)d_out);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
