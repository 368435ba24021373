// RoIAlign backward for sm_100a.
//
// Replaces ROIAlign_backward_cuda (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:428-487): at::zeros(B,C,H,W) followed by
// RoIAlignBackwardFeature (:259-365), which scatters every sample's four taps with fp32 atomicAdd (:340-347), once per
// FPN level per pooler.  Here ONE call produces the dense gradient of every level of the pyramid.
//
// CPM_BWD_DETERMINISTIC (NHWC fp32, fixed sampling grid): "pixel tiles own their gradient".
//   1. bwd_bin_rois: RoIs are binned by (level, image) in RoI-index order (ballot compaction; no atomics).
//   2. bwd_tiles: one CTA per 8x8-pixel tile x 128-channel chunk.  It scans the RoIs of its (level, image), and for
//      every RoI whose footprint reaches the tile it
//        - builds the two separable tap-weight tables  Ay[tile row][bin row], Ax[tile col][bin col]
//          (bilinear_interpolate_gradient :113-171 factors per axis: w1 = hy*hx, ...),
//        - stages the needed sub-rectangle of grad_out[r, c0:c0+128, :, :] through shared memory (transposing the
//          (channel, bin) order of the reference's output layout into channel-vector rows),
//        - accumulates  g[y][x][c] += Ay[y][p] * Ax[x][q] * go[r][c][p][q]  in registers, a warp per tile row, a lane
//          per 4 channels, RoIs in index order, bins in raster order -- a fixed summation order, no atomics.
//      Every pixel of every level is written exactly once (zeros where no RoI reaches), so no memset pass exists.
// CPM_BWD_ATOMIC: zero-fill + scatter; NHWC fp32 uses red.global.add.v4.f32 (one 16-byte reduction per lane per tap),
//   anything else (NCHW, fp64, nearest, adaptive grid) the reference-shaped scalar atomicAdd kernel.
#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);
int check_pyramid(const cpm_pyramid_t* p, const char* what);

template <typename T>
__device__ __forceinline__ int roi_level_b(const T* roi, const PyramidView& pv, const MapperView& mp, const int* lv,
                                           long n) {
  if (pv.num_levels == 1) return 0;
  if (lv) return lv[n];
  return fpn_level((float)roi[1], (float)roi[2], (float)roi[3], (float)roi[4], mp);
}

// ------------------------------------------------------------------------------------------------------------------
// generic atomic scatter (reference-shaped)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) roi_align_bwd_generic(PyramidView pv, const T* __restrict__ go,
                                                              const T* __restrict__ rois, long K, int PH, int PW, int sr,
                                                              int aligned, int interp, MapperView mp,
                                                              const int* __restrict__ roi_levels) {
  const int C = pv.channels;
  const long total = K * C * PH * PW;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int pw = idx % PW;
    int ph = (idx / PW) % PH;
    int c = (idx / PW / PH) % C;
    long n = idx / PW / PH / C;
    const T* roi = rois + 5 * n;
    int l = roi_level_b(roi, pv, mp, roi_levels, n);
    if (l < 0 || l >= pv.num_levels) continue;
    const int H = pv.H[l], W = pv.W[l];
    RoiGeo<T> g = roi_geometry<T>(roi, (T)pv.scale[l], PH, PW, sr, aligned != 0);
    if (g.b < 0 || g.b >= pv.batch) continue;
    long sC, sY, sX;
    if (pv.layout == CPM_LAYOUT_NCHW) {
      sC = (long)H * W; sY = W; sX = 1;
    } else {
      sC = 1; sY = (long)W * C; sX = C;
    }
    T* gi = (T*)pv.ptr[l] + (long)g.b * C * H * W + c * sC;
    const T top = go[idx];
    const T count = (T)(g.gh * g.gw);   // ROIAlign_cuda.cu:315 (no max(.,1) in the backward)
    for (int iy = 0; iy < g.gh; iy++) {
      const T y = g.start_h + ph * g.bin_h + static_cast<T>(iy + .5f) * g.bin_h / static_cast<T>(g.gh);
      for (int ix = 0; ix < g.gw; ix++) {
        const T x = g.start_w + pw * g.bin_w + static_cast<T>(ix + .5f) * g.bin_w / static_cast<T>(g.gw);
        if (interp == CPM_INTERP_BILINEAR) {
          // bilinear_interpolate_gradient, ROIAlign_cuda.cu:113-171
          T yy = y, xx = x;
          if (yy < (T)-1.0 || yy > (T)H || xx < (T)-1.0 || xx > (T)W) continue;
          if (yy <= 0) yy = 0;
          if (xx <= 0) xx = 0;
          int yl = (int)yy, xl = (int)xx, yh, xh;
          if (yl >= H - 1) { yh = yl = H - 1; yy = (T)yl; } else { yh = yl + 1; }
          if (xl >= W - 1) { xh = xl = W - 1; xx = (T)xl; } else { xh = xl + 1; }
          T ly = yy - yl, lx = xx - xl, hy = (T)1. - ly, hx = (T)1. - lx;
          T g1 = top * (hy * hx) / count, g2 = top * (hy * lx) / count;
          T g3 = top * (ly * hx) / count, g4 = top * (ly * lx) / count;
          atomicAdd(gi + yl * sY + xl * sX, g1);
          atomicAdd(gi + yl * sY + xh * sX, g2);
          atomicAdd(gi + yh * sY + xl * sX, g3);
          atomicAdd(gi + yh * sY + xh * sX, g4);
        } else {
          // nearest_interpolate_gradient, ROIAlign_cuda.cu:89-110
          if (y < (T)-0.5 || y >= (T)H - (T)0.5 || x < (T)-0.5 || x >= (T)W - (T)0.5) continue;
          int xl = (int)round(x), yl = (int)round(y);
          atomicAdd(gi + yl * sY + xl * sX, top / count);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NHWC vector-reduction scatter
// ------------------------------------------------------------------------------------------------------------------
constexpr int kChunk = 128;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ int stage_off(int c, int S) { return c * S + (c >> 5); }

__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_red(PyramidView pv, const float* __restrict__ go,
                                                               const float* __restrict__ rois, int PH, int PW, int sr,
                                                               int aligned, MapperView mp,
                                                               const int* __restrict__ roi_levels, int chunks, int S) {
  extern __shared__ float tile[];
  const int C = pv.channels;
  const long n = blockIdx.x / chunks;
  const int c0 = (blockIdx.x % chunks) * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cc = min(kChunk, C - c0);
  const float* roi = rois + 5 * n;
  const int l = roi_level_b(roi, pv, mp, roi_levels, n);
  if (l < 0 || l >= pv.num_levels) return;
  const int H = pv.H[l], W = pv.W[l];
  const RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, sr, aligned != 0);
  if (g.b < 0 || g.b >= pv.batch) return;
  const int PP = PH * PW;
  const float* src = go + ((long)n * C + c0) * PP;
  for (int e = threadIdx.x; e < cc * PP; e += blockDim.x) {
    const int c = e / PP, b = e - c * PP;
    tile[stage_off(c, S) + b] = __ldg(src + e);
  }
  __syncthreads();
  if (4 * lane >= cc) return;
  const float count = (float)(g.gh * g.gw);
  float* gi = (float*)pv.ptr[l] + (long)g.b * H * W * C + c0 + 4 * lane;
  const int c = 4 * lane;
  for (int bin = warp; bin < PP; bin += nwarps) {
    const int ph = bin / PW, pw = bin % PW;
    const float t0 = tile[stage_off(c + 0, S) + bin], t1 = tile[stage_off(c + 1, S) + bin];
    const float t2 = tile[stage_off(c + 2, S) + bin], t3 = tile[stage_off(c + 3, S) + bin];
    for (int iy = 0; iy < g.gh; iy++) {
      const float y = g.start_h + ph * g.bin_h + static_cast<float>(iy + .5f) * g.bin_h / static_cast<float>(g.gh);
      const AxisTap ty = axis_tap(y, H);
      if (!ty.valid) continue;
      for (int ix = 0; ix < g.gw; ix++) {
        const float x = g.start_w + pw * g.bin_w + static_cast<float>(ix + .5f) * g.bin_w / static_cast<float>(g.gw);
        const AxisTap tx = axis_tap(x, W);
        if (!tx.valid) continue;
        const float w1 = ty.wlo * tx.wlo / count, w2 = ty.wlo * tx.whi / count;
        const float w3 = ty.whi * tx.wlo / count, w4 = ty.whi * tx.whi / count;
        red_add_v4(gi + ((long)ty.lo * W + tx.lo) * C, t0 * w1, t1 * w1, t2 * w1, t3 * w1);
        red_add_v4(gi + ((long)ty.lo * W + tx.hi) * C, t0 * w2, t1 * w2, t2 * w2, t3 * w2);
        red_add_v4(gi + ((long)ty.hi * W + tx.lo) * C, t0 * w3, t1 * w3, t2 * w3, t3 * w3);
        red_add_v4(gi + ((long)ty.hi * W + tx.hi) * C, t0 * w4, t1 * w4, t2 * w4, t3 * w4);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// deterministic tile-owner gather
// ------------------------------------------------------------------------------------------------------------------
constexpr int TH = 8, TW = 8;          // tile: 8 rows (one per warp) x 8 columns (register-unrolled)
constexpr int kTileThreads = 256;
constexpr int kMaxP = 32;              // pooled size limit of this path (bin ranges are 32-bit ballots)
constexpr int kStageBins = 32;         // bins of grad_out staged per round

struct TileGrid {
  int tiles_x[CPM_MAX_LEVELS], tiles_y[CPM_MAX_LEVELS];
  int first[CPM_MAX_LEVELS + 1];       // first tile id of the level in launch order (coarsest level first)
  int order[CPM_MAX_LEVELS];           // launch order -> level
};

// workspace layout: int32 seg_count[L*B] ; int32 perm[L*B][K]
__global__ void __launch_bounds__(256) bwd_bin_rois(PyramidView pv, const float* __restrict__ rois, int K, MapperView mp,
                                                     const int* __restrict__ roi_levels, int* __restrict__ seg_count,
                                                     int* __restrict__ perm) {
  __shared__ int wcount[8];
  const int seg = blockIdx.x;            // seg = level * B + image
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = ((K + 7) / 8 + 31) & ~31;
  const int beg = warp * per_warp, end = min(K, beg + per_warp);
  int cnt = 0;
  unsigned long long mybits = 0;         // match flags of this lane's RoIs (<= 64 rounds kept; else recomputed)
  for (int i = beg + lane, it = 0; i - lane < end; i += 32, it++) {
    bool m = false;
    if (i < end) {
      const float* roi = rois + 5 * (long)i;
      const int l = roi_level_b(roi, pv, mp, roi_levels, i);
      const int b = (int)roi[0];
      m = l >= 0 && l < pv.num_levels && b >= 0 && b < pv.batch && l * pv.batch + b == seg;
    }
    if (it < 64 && m) mybits |= 1ull << it;
    cnt += __popc(__ballot_sync(0xffffffffu, m));
  }
  if (lane == 0) wcount[warp] = cnt;
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += wcount[w];
    total += wcount[w];
  }
  if (threadIdx.x == 0) seg_count[seg] = total;
  int* out = perm + (long)seg * K + base;
  for (int i = beg + lane, it = 0; i - lane < end; i += 32, it++) {
    bool m = false;
    if (it < 64) {
      m = (mybits >> it) & 1;
    } else if (i < end) {
      const float* roi = rois + 5 * (long)i;
      const int l = roi_level_b(roi, pv, mp, roi_levels, i);
      const int b = (int)roi[0];
      m = l >= 0 && l < pv.num_levels && b >= 0 && b < pv.batch && l * pv.batch + b == seg;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (m) out[__popc(bal & ((1u << lane) - 1))] = i;
    out += __popc(bal);
  }
}

struct Cand {
  int roi;
  float start_w, start_h, bin_w, bin_h;
};

// swizzled staging row: bin e, channel c (of the chunk) -> word index; conflict-free both for the (8 bins x 4 channels)
// patch stores and for the one-float4-per-lane row loads
__device__ __forceinline__ int sgo_off(int e, int c) { return e * kChunk + ((((c >> 2) ^ (e & 7))) << 2) + (c & 3); }

__global__ void __launch_bounds__(kTileThreads) bwd_tiles(PyramidView pv, TileGrid tg, const float* __restrict__ go,
                                                           const float* __restrict__ rois, int K, int PH, int PW, int G,
                                                           int aligned, const int* __restrict__ seg_count,
                                                           const int* __restrict__ perm, int chunks) {
  __shared__ Cand cand[kTileThreads];
  __shared__ float Ay[TH][kMaxP], Ax[TW][kMaxP];
  __shared__ unsigned pmask[TH], qmask[TW];
  __shared__ int wsum[8];
  __shared__ __align__(16) float sgo[kStageBins * kChunk];

  const int C = pv.channels;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- which tile ----
  int t = blockIdx.x / chunks;
  const int c0 = (blockIdx.x % chunks) * kChunk;
  int oi = 0;
  while (oi + 1 < pv.num_levels && t >= tg.first[oi + 1]) oi++;
  const int l = tg.order[oi];
  t -= tg.first[oi];
  const int per_img = tg.tiles_x[l] * tg.tiles_y[l];
  const int b = t / per_img;
  t -= b * per_img;
  const int y0 = (t / tg.tiles_x[l]) * TH, x0 = (t % tg.tiles_x[l]) * TW;
  const int H = pv.H[l], W = pv.W[l];
  const float scale = pv.scale[l];
  const int cc = min(kChunk, C - c0);
  const bool active = 4 * lane < cc;
  const int PP = PH * PW;
  const float invG = 1.0f / (float)G;

  float4 acc[TW];
#pragma unroll
  for (int x = 0; x < TW; x++) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int seg = l * pv.batch + b;
  const int nseg = seg_count[seg];
  const int* plist = perm + (long)seg * K;

  for (int base = 0; base < nseg; base += kTileThreads) {
    // ---- candidates of this round, in list order ----
    bool hit = false;
    Cand me;
    me.roi = -1;
    if (base + threadIdx.x < nseg) {
      me.roi = plist[base + threadIdx.x];
      const RoiGeo<float> g = roi_geometry<float>(rois + 5 * (long)me.roi, scale, PH, PW, G, aligned != 0);
      me.start_w = g.start_w; me.start_h = g.start_h; me.bin_w = g.bin_w; me.bin_h = g.bin_h;
      const float yf = g.start_h + static_cast<float>(.5f) * g.bin_h / static_cast<float>(G);
      const float yl = g.start_h + (PH - 1) * g.bin_h + static_cast<float>(G - 1 + .5f) * g.bin_h / static_cast<float>(G);
      const float xf = g.start_w + static_cast<float>(.5f) * g.bin_w / static_cast<float>(G);
      const float xl = g.start_w + (PW - 1) * g.bin_w + static_cast<float>(G - 1 + .5f) * g.bin_w / static_cast<float>(G);
      // rows/cols any tap can reach (conservative by construction: floor of the first sample .. floor of the last + 1)
      const bool none = yl < -1.0f || yf > (float)H || xl < -1.0f || xf > (float)W;
      const int ylo = yf <= 0.f ? 0 : min((int)yf, H - 1), yhi = yl >= (float)(H - 1) ? H - 1 : (int)fmaxf(yl, 0.f) + 1;
      const int xlo = xf <= 0.f ? 0 : min((int)xf, W - 1), xhi = xl >= (float)(W - 1) ? W - 1 : (int)fmaxf(xl, 0.f) + 1;
      hit = !none && ylo < y0 + TH && yhi >= y0 && xlo < x0 + TW && xhi >= x0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int pos = __popc(bal & ((1u << lane) - 1)), ncand = 0;
    for (int w = 0; w < 8; w++) {
      if (w < warp) pos += wsum[w];
      ncand += wsum[w];
    }
    if (hit) cand[pos] = me;
    __syncthreads();

    for (int ci = 0; ci < ncand; ci++) {
      const Cand cd = cand[ci];
      // ---- separable tap-weight tables ----
      for (int e = threadIdx.x; e < TH * PH + TW * PW; e += kTileThreads) {
        const bool isy = e < TH * PH;
        const int ee = isy ? e : e - TH * PH;
        const int P = isy ? PH : PW;
        const int r = ee / P, p = ee - r * P;
        const int pix = (isy ? y0 : x0) + r;
        const int size = isy ? H : W;
        const float start = isy ? cd.start_h : cd.start_w, bin = isy ? cd.bin_h : cd.bin_w;
        float w = 0.f;
        for (int i = 0; i < G; i++) {
          const float v = start + p * bin + static_cast<float>(i + .5f) * bin / static_cast<float>(G);
          const AxisTap tp = axis_tap(v, size);
          if (tp.valid) {
            if (tp.lo == pix) w += tp.wlo;
            if (tp.hi == pix) w += tp.whi;
          }
        }
        w *= invG;
        if (isy) Ay[r][p] = w; else Ax[r][p] = w;
      }
      __syncthreads();
      {
        const float wy = lane < PH ? Ay[warp][lane] : 0.f;
        const float wx = lane < PW ? Ax[warp][lane] : 0.f;
        const unsigned my = __ballot_sync(0xffffffffu, wy != 0.f), mx = __ballot_sync(0xffffffffu, wx != 0.f);
        if (lane == 0) { pmask[warp] = my; qmask[warp] = mx; }
      }
      __syncthreads();
      unsigned pall = 0, qall = 0;
#pragma unroll
      for (int i = 0; i < TH; i++) { pall |= pmask[i]; qall |= qmask[i]; }
      if (pall == 0 || qall == 0) continue;     // uniform: nothing of this RoI lands in the tile (barriers stay balanced: none below was entered)
      const int Pa = __ffs(pall) - 1, Pb = 32 - __clz(pall);
      const int Qa = __ffs(qall) - 1, Qb = 32 - __clz(qall);
      const int nq = Qb - Qa;
      const int rows_per_round = kStageBins / nq;       // nq <= 32
      const unsigned myp = pmask[warp];
      unsigned qm[TW];
#pragma unroll
      for (int x = 0; x < TW; x++) qm[x] = qmask[x];
      const float* gsrc = go + ((long)cd.roi * C + c0) * PP;

      for (int pr = Pa; pr < Pb; pr += rows_per_round) {
        const int np = min(rows_per_round, Pb - pr);
        const int nb = np * nq;
        // ---- stage grad_out[r, c0:c0+cc, pr:pr+np, Qa:Qb] -> sgo[bin][channel] ----
        {
          const int j = lane & 3, eb = lane >> 2;
          for (int e0 = 0; e0 < nb; e0 += 8) {
            const int e = e0 + eb;
            if (e < nb) {
              const int pe = e / nq;
              const float* src = gsrc + (pr + pe) * PW + Qa + (e - pe * nq);
              for (int cg = warp; 4 * cg < cc; cg += 8) {
                const int c = 4 * cg + j;
                sgo[sgo_off(e, c)] = __ldg(src + (long)c * PP);
              }
            }
          }
        }
        __syncthreads();
        // ---- gather ----
        if (active) {
          unsigned pm = myp & (np >= 32 ? 0xffffffffu : (((1u << np) - 1) << pr));
          while (pm) {
            const int p = __ffs(pm) - 1;
            pm &= pm - 1;
            const float wy = Ay[warp][p];
            const int erow = (p - pr) * nq - Qa;
#pragma unroll
            for (int x = 0; x < TW; x++) {
              unsigned m = qm[x];
              while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                const float w = wy * Ax[x][q];
                const int e = erow + q;
                const float4 v = *reinterpret_cast<const float4*>(&sgo[e * kChunk + ((lane ^ (e & 7)) << 2)]);
                acc[x].x = fmaf(w, v.x, acc[x].x);
                acc[x].y = fmaf(w, v.y, acc[x].y);
                acc[x].z = fmaf(w, v.z, acc[x].z);
                acc[x].w = fmaf(w, v.w, acc[x].w);
              }
            }
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }

  // ---- the tile's gradient: written exactly once ----
  const int y = y0 + warp;
  if (active && y < H) {
    float4* dst = reinterpret_cast<float4*>((float*)pv.ptr[l] + (((long)b * H + y) * W + x0) * C + c0) + lane;
    const long C4 = C >> 2;
#pragma unroll
    for (int x = 0; x < TW; x++)
      if (x0 + x < W) dst[x * C4] = acc[x];
  }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace cpm

using namespace cpm;

extern "C" size_t cpm_roi_align_backward_workspace_bytes(int64_t K, int num_levels, int batch) {
  const size_t segs = (size_t)(num_levels > 0 ? num_levels : 1) * (size_t)(batch > 0 ? batch : 1);
  return align256(segs * sizeof(int)) + align256(segs * (size_t)(K > 0 ? K : 1) * sizeof(int));
}

extern "C" int cpm_roi_align_backward(const cpm_pyramid_t* grad_feat, const void* d_grad_out, const void* d_rois, int64_t K,
                                      int pooled_h, int pooled_w, int sampling_ratio, int aligned, int interpolation,
                                      const cpm_level_mapper_t* mapper, const int32_t* d_roi_levels, int mode,
                                      void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_pyramid(grad_feat, "grad_feat");
  if (rc != CPM_OK) return rc;
  CPM_CHECK_ARG(K >= 0, "K < 0");
  CPM_CHECK_ARG(K < (1L << 30), "K too large");
  CPM_CHECK_ARG(pooled_h >= 1 && pooled_w >= 1, "pooled size must be positive");
  CPM_CHECK_ARG(interpolation == CPM_INTERP_BILINEAR || interpolation == CPM_INTERP_NEAREST,
                "unknown interpolation method %d", interpolation);
  CPM_CHECK_ARG(mode == CPM_BWD_DETERMINISTIC || mode == CPM_BWD_ATOMIC, "unknown backward mode %d", mode);
  CPM_CHECK_ARG(grad_feat->num_levels == 1 || mapper != nullptr || d_roi_levels != nullptr,
                "a multi-level pyramid needs a level mapper or per-RoI levels");
  if (grad_feat->dtype == CPM_BF16) {
    set_error("bf16 gradients are not supported: keep the feature gradient in fp32");
    return CPM_ERR_UNSUPPORTED;
  }
  if (K > 0) {
    if ((rc = check_device_ptr(d_rois, "rois")) != CPM_OK) return rc;
    if ((rc = check_device_ptr(d_grad_out, "grad_out")) != CPM_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PyramidView pv = make_view(grad_feat);
  MapperView mp = make_view(mapper);
  const int C = grad_feat->channels, B = grad_feat->batch, L = grad_feat->num_levels;
  const size_t esz = grad_feat->dtype == CPM_F64 ? 8 : 4;
  if (B == 0) return CPM_OK;

  bool nhwc_f32 = grad_feat->layout == CPM_LAYOUT_NHWC && grad_feat->dtype == CPM_F32 &&
                  interpolation == CPM_INTERP_BILINEAR && C % 4 == 0;
  for (int l = 0; nhwc_f32 && l < L; l++) nhwc_f32 = ((uintptr_t)grad_feat->d_level[l] & 15) == 0;
  const int chunks = (C + kChunk - 1) / kChunk;

  if (mode == CPM_BWD_DETERMINISTIC) {
    if (!(nhwc_f32 && sampling_ratio >= 1 && pooled_h <= kMaxP && pooled_w <= kMaxP)) {
      set_error("deterministic backward needs an NHWC fp32 gradient pyramid, bilinear interpolation, C %% 4 == 0, "
                "sampling_ratio >= 1 and pooled size <= %d; use CPM_BWD_ATOMIC otherwise", kMaxP);
      return CPM_ERR_UNSUPPORTED;
    }
    const size_t need = cpm_roi_align_backward_workspace_bytes(K, L, B);
    if (d_workspace == nullptr || workspace_bytes < need) {
      set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need);
      return CPM_ERR_WORKSPACE;
    }
    if ((rc = check_device_ptr(d_workspace, "workspace")) != CPM_OK) return rc;
    int* seg_count = (int*)d_workspace;
    int* perm = (int*)((char*)d_workspace + align256((size_t)L * B * sizeof(int)));
    const int Kp = K > 0 ? (int)K : 1;
    if (K > 0) {
      bwd_bin_rois<<<L * B, 256, 0, st>>>(pv, (const float*)d_rois, (int)K, mp, d_roi_levels, seg_count, perm);
      CPM_CHECK_LAUNCH();
    } else {
      CPM_CHECK_CUDA(cudaMemsetAsync(seg_count, 0, (size_t)L * B * sizeof(int), st));
    }
    TileGrid tg;
    long tiles = 0;
    for (int i = 0; i < L; i++) {
      const int l = L - 1 - i;   // coarse levels (many RoIs per tile) first
      tg.order[i] = l;
      tg.tiles_x[l] = (grad_feat->width[l] + TW - 1) / TW;
      tg.tiles_y[l] = (grad_feat->height[l] + TH - 1) / TH;
      tg.first[i] = (int)tiles;
      tiles += (long)B * tg.tiles_x[l] * tg.tiles_y[l];
    }
    for (int i = L; i <= CPM_MAX_LEVELS; i++) tg.first[i] = (int)tiles;
    CPM_CHECK_ARG(tiles * chunks < (1L << 31), "gradient pyramid too large for one launch");
    bwd_tiles<<<(unsigned)(tiles * chunks), kTileThreads, 0, st>>>(pv, tg, (const float*)d_grad_out, (const float*)d_rois,
                                                                  Kp, pooled_h, pooled_w, sampling_ratio, aligned,
                                                                  seg_count, perm, chunks);
    CPM_CHECK_LAUNCH();
    return CPM_OK;
  }

  // ---- atomic mode: zero-fill (ROIAlign_cuda.cu:451-452) + scatter ----
  for (int l = 0; l < L; l++)
    CPM_CHECK_CUDA(cudaMemsetAsync(grad_feat->d_level[l], 0,
                                   (size_t)B * C * grad_feat->height[l] * grad_feat->width[l] * esz, st));
  if (K == 0) return CPM_OK;
  const int PP = pooled_h * pooled_w;
  const size_t smem = (size_t)(kChunk * (PP | 1) + 8) * sizeof(float);
  if (nhwc_f32 && smem <= 200 * 1024 && (long)K * chunks < (1L << 31)) {
    static thread_local int configured_dev = -1;
    int dev;
    CPM_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      CPM_CHECK_CUDA(
          cudaFuncSetAttribute(roi_align_bwd_nhwc_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured_dev = dev;
    }
    roi_align_bwd_nhwc_red<<<(unsigned)(K * chunks), 256, smem, st>>>(pv, (const float*)d_grad_out, (const float*)d_rois,
                                                                      pooled_h, pooled_w, sampling_ratio, aligned, mp,
                                                                      d_roi_levels, chunks, PP | 1);
    CPM_CHECK_LAUNCH();
    return CPM_OK;
  }
  const long total = (long)K * C * PP;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 64) blocks = 148L * 64;
  if (grad_feat->dtype == CPM_F32)
    roi_align_bwd_generic<float><<<(unsigned)blocks, 256, 0, st>>>(pv, (const float*)d_grad_out, (const float*)d_rois, K,
                                                                   pooled_h, pooled_w, sampling_ratio, aligned,
                                                                   interpolation, mp, d_roi_levels);
  else
    roi_align_bwd_generic<double><<<(unsigned)blocks, 256, 0, st>>>(pv, (const double*)d_grad_out, (const double*)d_rois,
                                                                    K, pooled_h, pooled_w, sampling_ratio, aligned,
                                                                    interpolation, mp, d_roi_levels);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
