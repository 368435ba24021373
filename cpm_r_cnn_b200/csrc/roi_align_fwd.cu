// RoIAlign forward for sm_100a.
//
// Replaces ROIAlign_forward_cuda (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:367-425, kernel :178-256) and, for a
// multi-level pyramid, the per-level nonzero/gather/index_put loop of Pooler.forward (pet/rcnn/utils/poolers.py:117-131).
//
// Two kernels:
//   * roi_align_fwd_generic<T>: one thread per output element, any layout / dtype / interpolation / sampling grid.
//     The arithmetic is the reference kernel's, statement by statement.
//   * roi_align_fwd_nhwc<...>: the hot path.  One CTA per (RoI, 128-channel chunk); a warp owns whole rows of output
//     bins, a lane owns 4 consecutive channels, so every bilinear tap is one coalesced 512-byte row of float4 loads
//     from the NHWC map.  The pooled block is transposed through shared memory and leaves as one contiguous,
//     16-byte-vectorised stream into the (K, C, PH, PW) output -- the layout the reference returns.
//     The FPN level of each RoI (LevelMapper, poolers.py:29-40) is evaluated inside the kernel.
#include <cuda_bf16.h>

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
constexpr int kMaxRowsPH = 32;       // pooled height limit of the register-row kernel

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
  if (l >= 0 && l < pv.num_levels) {
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

// Compile-time-shaped kernel (the CPM head's 7x7 and 14x14 poolers with sampling_ratio 2, config.py:881-886):
// a warp owns one row of PW bins and keeps all PW accumulators (4 channels each) in registers while it sweeps the
// row's G*PW sample columns left to right, G sample rows one after the other -- the reference's summation order
// (iy outer, ix inner, ROIAlign_cuda.cu:234-252).  Two neighbouring sample columns usually share a feature column
// (x_high of one is x_low of the next); the sweep keeps the last two columns of both feature rows in registers and
// only loads a column when the (warp-uniform) index changes, which halves the L1 traffic of the naive 4-tap gather.
struct __align__(16) TapS {
  int lo, hi;       // lo < 0: sample out of range (contributes 0, ROIAlign_cuda.cu:46-49)
  float wlo, whi;
};

__device__ __forceinline__ TapS make_tap(float v, int size) {
  const AxisTap t = axis_tap(v, size);
  TapS r;
  r.lo = t.valid ? t.lo : -1;
  r.hi = t.hi;
  r.wlo = t.wlo;
  r.whi = t.whi;
  return r;
}

template <int PW, int G>
__global__ void __launch_bounds__(32 * 7, 2) roi_align_fwd_nhwc_rows(PyramidView pv, const float* __restrict__ rois, int PH,
                                                                      int aligned, MapperView mp,
                                                                      const int* __restrict__ roi_levels,
                                                                      float* __restrict__ out, int chunks, int S) {
  extern __shared__ float tile[];
  __shared__ TapS xt[PW * G];
  __shared__ TapS yt[kMaxRowsPH * G];
  const int C = pv.channels;
  const long n = blockIdx.x / chunks;
  const int c0 = (blockIdx.x % chunks) * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cc = min(kChunk, C - c0);
  const float* roi = rois + 5 * n;
  const int l = roi_level(roi, pv, mp, roi_levels, n);
  const int PP = PH * PW;
  if (l >= 0 && l < pv.num_levels) {
    const int H = pv.H[l], W = pv.W[l];
    const RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, G, aligned != 0);
    // the RoI's sample taps along each axis, once per CTA (every bin row shares the x taps, every bin column the y taps)
    for (int e = threadIdx.x; e < (PW + PH) * G; e += blockDim.x) {
      const bool isx = e < PW * G;
      const int k = isx ? e : e - PW * G;
      const int p = k / G, i = k % G;
      const float start = isx ? g.start_w : g.start_h, bin = isx ? g.bin_w : g.bin_h;
      const float v = start + p * bin + static_cast<float>(i + .5f) * bin / static_cast<float>(G);
      if (isx) xt[k] = make_tap(v, W); else yt[k] = make_tap(v, H);
    }
    __syncthreads();
    const float count = (float)(G * G);
    const bool active = 4 * lane < cc;
    const float4* f = (const float4*)((const float*)pv.ptr[l] + (long)g.b * H * W * C + c0) + lane;
    const long C4 = C >> 2;
    for (int ph = warp; ph < PH; ph += nwarps) {
      float4 acc[PW];
#pragma unroll
      for (int i = 0; i < PW; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
#pragma unroll 1
        for (int iy = 0; iy < G; iy++) {
          const TapS ty = yt[ph * G + iy];
          if (ty.lo < 0) continue;
          const float4* rlo = f + (long)ty.lo * W * C4;
          const float4* rhi = f + (long)ty.hi * W * C4;
          int cx0 = -1, cx1 = -1;       // feature columns currently held: (a0,b0) = column cx0 of rows lo/hi, (a1,b1) = cx1
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), b0 = a0, a1 = a0, b1 = a0;
#pragma unroll
          for (int sx = 0; sx < PW * G; sx++) {
            const int pw = sx / G;
            const TapS tx = xt[sx];
            if (tx.lo < 0) continue;
            if (tx.lo != cx0) {
              if (tx.lo == cx1) {
                a0 = a1; b0 = b1;
              } else {
                a0 = ldg_f4(rlo + (long)tx.lo * C4);
                b0 = ldg_f4(rhi + (long)tx.lo * C4);
              }
              cx0 = tx.lo;
            }
            if (tx.hi != cx1) {
              if (tx.hi == cx0) {
                a1 = a0; b1 = b0;
              } else {
                a1 = ldg_f4(rlo + (long)tx.hi * C4);
                b1 = ldg_f4(rhi + (long)tx.hi * C4);
              }
              cx1 = tx.hi;
            }
            const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
            acc[pw].x += tap4(w1, a0.x, w2, a1.x, w3, b0.x, w4, b1.x);
            acc[pw].y += tap4(w1, a0.y, w2, a1.y, w3, b0.y, w4, b1.y);
            acc[pw].z += tap4(w1, a0.z, w2, a1.z, w3, b0.z, w4, b1.z);
            acc[pw].w += tap4(w1, a0.w, w2, a1.w, w3, b0.w, w4, b1.w);
          }
        }
        const int c = 4 * lane;
        float* t0 = tile + stage_off(c + 0, S) + ph * PW;
        float* t1 = tile + stage_off(c + 1, S) + ph * PW;
        float* t2 = tile + stage_off(c + 2, S) + ph * PW;
        float* t3 = tile + stage_off(c + 3, S) + ph * PW;
#pragma unroll
        for (int pw = 0; pw < PW; pw++) {
          t0[pw] = acc[pw].x / count;
          t1[pw] = acc[pw].y / count;
          t2[pw] = acc[pw].z / count;
          t3[pw] = acc[pw].w / count;
        }
      }
    }
  } else {
    for (int e = threadIdx.x; e < cc * PP; e += blockDim.x) tile[stage_off(e / PP, S) + e % PP] = 0.f;
  }
  __syncthreads();
  float* o = out + ((long)n * C + c0) * PP;
  const int total = cc * PP;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e / PP, b = e - c * PP;
    o[e] = tile[stage_off(c, S) + b];
  }
}

template <int PW, int G>
static int launch_rows(const PyramidView& pv, const float* rois, long K, int PH, int aligned, const MapperView& mp,
                       const int* lv, float* out, cudaStream_t st) {
  const int chunks = (pv.channels + kChunk - 1) / kChunk;
  const int PP = PH * PW;
  const int S = PP | 1;
  const size_t smem = (size_t)(kChunk * S + 8) * sizeof(float);
  auto kern = roi_align_fwd_nhwc_rows<PW, G>;
  static thread_local int configured_dev = -1;
  int dev;
  CPM_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  const int warps = PH < 7 ? PH : 7;
  kern<<<(unsigned)(K * chunks), 32 * warps, smem, st>>>(pv, rois, PH, aligned, mp, lv, out, chunks, S);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

int check_device_ptr(const void* p, const char* what);

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
  int rc = check_pyramid(feat, "feat");
  if (rc != CPM_OK) return rc;
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
  if (impl == CPM_FWD_NHWC && !nhwc_ok) {
    set_error("CPM_FWD_NHWC needs an NHWC fp32 pyramid, bilinear interpolation, C %% 4 == 0 and 16-byte aligned maps");
    return CPM_ERR_UNSUPPORTED;
  }
  if (nhwc_ok && impl != CPM_FWD_GENERIC) {
    if (sampling_ratio == 2 && pooled_w == 7 && pooled_h <= kMaxRowsPH)
      return launch_rows<7, 2>(pv, (const float*)d_rois, K, pooled_h, aligned, mp, d_roi_levels, (float*)d_out, st);
    if (sampling_ratio == 2 && pooled_w == 14 && pooled_h <= kMaxRowsPH)
      return launch_rows<14, 2>(pv, (const float*)d_rois, K, pooled_h, aligned, mp, d_roi_levels, (float*)d_out, st);
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
    set_error("bf16 pyramids need the NHWC kernel (layout NHWC, bilinear, C %% 8 == 0)");
    return CPM_ERR_UNSUPPORTED;
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
