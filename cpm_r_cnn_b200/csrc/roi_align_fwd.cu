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

// ------------------------------------------------------------------------------------------------------------------
// Column-sweep kernel (sampling_ratio 2; any pooled size up to 32x32): the hot path of the CPM head's poolers.
//
// One CTA = (RoI, 128-channel chunk, group of <= 7 bin rows); a warp owns one bin row, a lane 4 channels.  The warp walks
// the feature COLUMNS its row touches from left to right.  Per step it issues the loads of the next kColBatch columns
// of the four feature rows of its two sample rows (y_low/y_high of iy = 0, 1) as one unconditional batch -- 8
// independent 512-byte row loads in flight per warp -- and then consumes every sample whose x_low is one of those
// columns (a warp-uniform walk over the RoI's x-tap table in shared memory).  Every feature pixel of the band is
// loaded exactly once per bin row; both sample rows and all sample columns that share it are served from registers.
// The four samples of a bin are summed in the reference's order ((v00 + v01) + v10) + v11 (ROIAlign_cuda.cu:234-252).
// Finished bins go to a swizzled shared-memory tile [bin][channel]; the tile leaves as (4 channels x 8 bins) patches:
// bank-conflict-free reads, 32-byte-segment global writes into the reference's (K, C, PH, PW) layout.
constexpr int kRowsPerCta = 7;
constexpr int kColBatch = 2;

typedef unsigned long long u64;

// packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2): two IEEE-rounded fp32 operations per instruction, bit-identical
// to the scalar forms -- halves the issue slots of the bilinear math
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

struct V4 {          // 4 channels as two fp32x2 pairs
  u64 lo, hi;
};
struct Col4 {
  V4 r[4];
};
struct __align__(16) W8 {      // the four bilinear weights of one (sample row, sample column), each duplicated for fp32x2
  u64 w1, w2, w3, w4;
};

__device__ __forceinline__ V4 ldg_v4(const char* p) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
  V4 r;
  r.lo = v.x;
  r.hi = v.y;
  return r;
}

// w1*v1 + w2*v2 + w3*v3 + w4*v4, reference order, contracted as mul, fma, fma, fma (see tap4)
__device__ __forceinline__ V4 bilin4(const W8& w, const V4& v1, const V4& v2, const V4& v3, const V4& v4) {
  V4 r;
  r.lo = fma2(w.w4, v4.lo, fma2(w.w3, v3.lo, fma2(w.w2, v2.lo, mul2(w.w1, v1.lo))));
  r.hi = fma2(w.w4, v4.hi, fma2(w.w3, v3.hi, fma2(w.w2, v2.hi, mul2(w.w1, v1.hi))));
  return r;
}
__device__ __forceinline__ V4 add4(const V4& a, const V4& b) {
  V4 r;
  r.lo = add2(a.lo, b.lo);
  r.hi = add2(a.hi, b.hi);
  return r;
}

__global__ void __launch_bounds__(32 * kRowsPerCta, 2) roi_align_fwd_nhwc_sweep(
    PyramidView pv, const float* __restrict__ rois, int PH, int PW, int aligned, MapperView mp,
    const int* __restrict__ roi_levels, float* __restrict__ out, int chunks, int rgroups) {
  constexpr int G = 2;
  constexpr int kMaxS = kMaxRowsPH * G;
  extern __shared__ __align__(16) float smem_dyn[];
  // dynamic: W8 wtab[nrows*G][NS] | float tile[kChunk][nbp]
  __shared__ int xlo[kMaxS];
  __shared__ TapS xt[kMaxS];
  __shared__ TapS yt[kRowsPerCta * G];
  __shared__ int s_range[2];
  const int C = pv.channels;
  const int rg = blockIdx.x % rgroups;
  const int chunk = (blockIdx.x / rgroups) % chunks;
  const long n = blockIdx.x / (rgroups * chunks);
  const int c0 = chunk * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cc = min(kChunk, C - c0);
  const int ph0 = rg * kRowsPerCta;
  const int nrows = min(kRowsPerCta, PH - ph0);
  const int NS = PW * G;
  const int nb = nrows * PW;
  const int nbp = (kRowsPerCta * PW) | 1;             // odd row stride of the staging tile
  W8* wtab = reinterpret_cast<W8*>(smem_dyn);
  float* tile = smem_dyn + (size_t)kRowsPerCta * G * NS * (sizeof(W8) / sizeof(float));
  const float* roi = rois + 5 * n;
  const int l = roi_level(roi, pv, mp, roi_levels, n);
  const bool lvl_ok = l >= 0 && l < pv.num_levels;
  if (lvl_ok) {
    const int H = pv.H[l], W = pv.W[l];
    const RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, G, aligned != 0);
    // ---- per-CTA tables: axis taps, then the 4 weights of every (sample row, sample column) ----
    for (int e = threadIdx.x; e < NS + nrows * G; e += blockDim.x) {
      const bool isx = e < NS;
      const int k = isx ? e : e - NS + ph0 * G;
      const int p = k / G, i = k % G;
      const float start = isx ? g.start_w : g.start_h, bin = isx ? g.bin_w : g.bin_h;
      const float v = start + p * bin + static_cast<float>(i + .5f) * bin / static_cast<float>(G);
      if (isx) {
        const TapS t = make_tap(v, W);
        xt[k] = t;
        xlo[k] = t.lo;
      } else {
        yt[k - ph0 * G] = make_tap(v, H);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {          // valid samples are one contiguous run (coordinates are monotone)
      int sb = 0, se = NS;
      while (sb < NS && xlo[sb] < 0) sb++;
      while (se > sb && xlo[se - 1] < 0) se--;
      s_range[0] = sb;
      s_range[1] = se;
    }
    for (int e = threadIdx.x; e < nrows * G * NS; e += blockDim.x) {
      const TapS ty = yt[e / NS], tx = xt[e % NS];
      const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
      W8 w;
      w.w1 = pack2(w1, w1); w.w2 = pack2(w2, w2); w.w3 = pack2(w3, w3); w.w4 = pack2(w4, w4);
      wtab[e] = w;
    }
    __syncthreads();
    if (warp < nrows && 4 * lane < cc) {
      const TapS ty0 = yt[warp * G], ty1 = yt[warp * G + 1];
      const bool yv0 = ty0.lo >= 0, yv1 = ty1.lo >= 0;
      const size_t stride = (size_t)C * sizeof(float);                 // bytes per pixel
      const char* f = reinterpret_cast<const char*>((const float*)pv.ptr[l] + (long)g.b * H * W * C + c0 + 4 * lane);
      const char* p0 = f + (size_t)(yv0 ? ty0.lo : 0) * W * stride;
      const char* p1 = f + (size_t)(yv0 ? ty0.hi : 0) * W * stride;
      const char* p2 = f + (size_t)(yv1 ? ty1.lo : 0) * W * stride;
      const char* p3 = f + (size_t)(yv1 ? ty1.hi : 0) * W * stride;
      const W8* w0 = wtab + (warp * G) * NS;
      const W8* w1t = w0 + NS;
      float* trow = tile + warp * PW;
      const int cch = 4 * lane;
      float* t0 = trow + (cch + 0) * nbp + (cch >> 5);
      float* t1p = trow + (cch + 1) * nbp + (cch >> 5);
      float* t2 = trow + (cch + 2) * nbp + (cch >> 5);
      float* t3 = trow + (cch + 3) * nbp + (cch >> 5);
      const int s_begin = s_range[0], s_end = s_range[1];
      V4 s0, t1;
      s0.lo = s0.hi = t1.lo = t1.hi = 0ull;            // +0.0f pairs
      const V4 zero4 = s0;

      auto load_col = [&](int x, int xmax) {
        const size_t xo = (size_t)min(x, xmax) * stride;
        Col4 c;
        c.r[0] = ldg_v4(p0 + xo);
        c.r[1] = ldg_v4(p1 + xo);
        c.r[2] = ldg_v4(p2 + xo);
        c.r[3] = ldg_v4(p3 + xo);
        return c;
      };
      auto flush = [&](int pw, const V4& v) {
        const float2 a = unpack2(v.lo), b = unpack2(v.hi);
        t0[pw] = a.x * 0.25f;
        t1p[pw] = a.y * 0.25f;
        t2[pw] = b.x * 0.25f;
        t3[pw] = b.y * 0.25f;
      };
      auto skip_sample = [&](int sx) {        // out-of-range sample: adds 0 (ROIAlign_cuda.cu:46-49)
        if (sx & 1) flush(sx >> 1, add4(s0, t1)); else { s0 = zero4; t1 = zero4; }
      };
      auto take_sample = [&](int sx, const Col4& ca, const Col4& cb) {
        V4 va = zero4, vb = zero4;
        if (yv0) va = bilin4(w0[sx], ca.r[0], cb.r[0], ca.r[1], cb.r[1]);
        if (yv1) vb = bilin4(w1t[sx], ca.r[2], cb.r[2], ca.r[3], cb.r[3]);
        if (sx & 1) {
          flush(sx >> 1, add4(add4(add4(s0, va), t1), vb));    // ((v00 + v01) + v10) + v11
        } else {
          s0 = va;
          t1 = vb;
        }
      };

      int s = 0;
      for (; s < s_begin; s++) skip_sample(s);
      if (s < s_end) {
        const int xmax = xt[s_end - 1].hi;         // right-most column any sample needs
        int col = xlo[s];
        Col4 c[kColBatch + 1];
        c[0] = load_col(col, xmax);
        while (s < s_end) {
#pragma unroll
          for (int k = 1; k <= kColBatch; k++) c[k] = load_col(col + k, xmax);
#pragma unroll
          for (int k = 0; k < kColBatch; k++) {
            while (s < s_end && xlo[s] == col + k) {
              take_sample(s, c[k], c[k + 1]);
              s++;
            }
          }
          c[0] = c[kColBatch];
          col += kColBatch;
          if (s < s_end) {
            const int nl = xlo[s];
            if (nl > col) {          // sample spacing > 2 pixels: jump over untouched columns
              col = nl;
              c[0] = load_col(col, xmax);
            }
          }
        }
      }
      for (; s < NS; s++) skip_sample(s);
    }
  } else {
    for (int e = threadIdx.x; e < kChunk * nbp + 4; e += blockDim.x) tile[e] = 0.f;
  }
  __syncthreads();
  // ---- tile -> out[n, c0:c0+cc, ph0:ph0+nrows, :]  (nb contiguous floats per channel) ----
  {
    const int PP = PH * PW;
    float* o = out + ((long)n * C + c0) * PP + ph0 * PW;
    for (int c = warp; c < cc; c += kRowsPerCta) {
      const float* tr = tile + c * nbp + (c >> 5);
      float* oc = o + (long)c * PP;
      for (int e = lane; e < nb; e += 32) oc[e] = tr[e];
    }
  }
}

static size_t sweep_smem_bytes(int PW) {
  const int NS = PW * 2;
  return (size_t)kRowsPerCta * 2 * NS * sizeof(W8) + (size_t)(kChunk * ((kRowsPerCta * PW) | 1) + 8) * sizeof(float);
}

static int launch_sweep(const PyramidView& pv, const float* rois, long K, int PH, int PW, int aligned, const MapperView& mp,
                        const int* lv, float* out, cudaStream_t st) {
  const int chunks = (pv.channels + kChunk - 1) / kChunk;
  const int rgroups = (PH + kRowsPerCta - 1) / kRowsPerCta;
  const size_t smem = sweep_smem_bytes(PW);
  static thread_local int configured_dev = -1;
  int dev;
  CPM_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CPM_CHECK_CUDA(cudaFuncSetAttribute(roi_align_fwd_nhwc_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  const long blocks = K * chunks * rgroups;
  CPM_CHECK_ARG(blocks < (1L << 31), "too many RoIs for one launch");
  roi_align_fwd_nhwc_sweep<<<(unsigned)blocks, 32 * kRowsPerCta, smem, st>>>(pv, rois, PH, PW, aligned, mp, lv, out, chunks,
                                                                            rgroups);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
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
bool fwd_cols_supported(const cpm_pyramid_t* feat, int pooled_h, int pooled_w, int sampling_ratio, const void* d_out);
int launch_fwd_cols(const PyramidView& pv, const float* rois, long K, int P, int G, int aligned, const MapperView& mp,
                    const int* lv, void* out, int out_channels_last, cudaStream_t st);

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
  if ((impl == CPM_FWD_NHWC || impl == CPM_FWD_NHWC_ROWS) && !nhwc_ok) {
    set_error("CPM_FWD_NHWC needs an NHWC fp32 pyramid, bilinear interpolation, C %% 4 == 0 and 16-byte aligned maps");
    return CPM_ERR_UNSUPPORTED;
  }
  if (nhwc_ok && impl != CPM_FWD_GENERIC) {
    bool small_maps = true;     // the sweep kernel indexes a level with 32-bit float4 offsets
    for (int l = 0; l < feat->num_levels; l++)
      small_maps = small_maps && (double)feat->batch * feat->height[l] * feat->width[l] * feat->channels < 8.0e9;
    if (sampling_ratio == 2 && pooled_w <= kMaxRowsPH && pooled_h <= kMaxRowsPH && small_maps && impl != CPM_FWD_NHWC_ROWS)
      return launch_sweep(pv, (const float*)d_rois, K, pooled_h, pooled_w, aligned, mp, d_roi_levels, (float*)d_out, st);
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
    CPM_CHECK_CUDA(cudaFuncSetAttribute(roi_align_fwd_nhwc_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
