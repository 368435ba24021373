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
#include <stdlib.h>
#include <string.h>

#include "roi_align_bwd.cuh"

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
constexpr int kCandCap = 64;           // prepared candidates per 8x8 tile; a busier tile is scanned by its CTAs

// workspace layout: int32 seg_count[L*B] ; int32 perm[L*B][K]
__device__ __forceinline__ void bin_rois_block(const PyramidView& pv, const float* __restrict__ rois, int K, const MapperView& mp,
                                               const int* __restrict__ roi_levels, int* __restrict__ seg_count,
                                               int* __restrict__ perm, const int seg) {
  __shared__ int wcount[8];              // seg = level * B + image
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

// 64 threads per RoI: warp 0 builds the y sample taps, warp 1 the x sample taps (P * G <= 32 per axis on this path), plus
// what the tile kernels need to start from a list instead of a scan:
//   taps    [K][(PH + PW) * G]  (y taps first)
//   box     [K] int4 {ylo, yhi, xlo, xhi}: the pixels the taps reach, ylo > yhi when nothing is reached
//   segid   [K] level * B + image, -1 for an RoI outside the pyramid / batch
//   rowclip [K][NBy], colclip [K][NBx]: per band of 8 pixel rows (columns) first | last << 8 bin row (column) with a
//           sample tap inside the band, 1 (first 1 > last 0) when there is none
__device__ __forceinline__ void roi_taps_group(const PyramidView& pv, const float* __restrict__ rois, int PH, int PW, int G,
                                               int aligned, const MapperView& mp, const int* __restrict__ roi_levels,
                                               TapS* __restrict__ taps, int4* __restrict__ box, const int r, const int tid,
                                               int* __restrict__ segid, int* __restrict__ rowclip, const int NBy,
                                               int* __restrict__ colclip, const int NBx, const TileGrid& tg8,
                                               int* __restrict__ tile_count, int2* __restrict__ cands) {
  // per group of 64 threads: the axis clips of the RoI's bands, exchanged between its two warps for the tile lists
  __shared__ int s_clip[4][2][64];
  __shared__ int s_b0[4][2], s_nb[4][2];
  const int grp = (threadIdx.x >> 6) & 3;
  const float* roi = rois + 5 * (long)r;
  const int l = roi_level_b(roi, pv, mp, roi_levels, r);
  const int nt = (PH + PW) * G;
  TapS* out = taps + (long)r * nt;
  const bool ok = l >= 0 && l < pv.num_levels && (int)roi[0] >= 0 && (int)roi[0] < pv.batch;
  const int isx = tid >> 5, ln = tid & 31;             // warp 0: y axis, warp 1: x axis
  int* clip = isx ? colclip : rowclip;
  const int NB = isx ? NBx : NBy;
  if (clip != nullptr)
    for (int b = ln; b < NB; b += 32) clip[(long)r * NB + b] = 1;
  if (!ok) {
    if (tid == 0) {
      box[r] = make_int4(1, 0, 1, 0);
      if (segid) segid[r] = -1;
    }
    return;
  }
  if (ln == 0) s_nb[grp][isx] = 0;
  const int H = pv.H[l], W = pv.W[l];
  const RoiGeo<float> g = roi_geometry<float>(roi, pv.scale[l], PH, PW, G, aligned != 0);
  const int P = isx ? PW : PH;
  TapS mine;
  mine.lo = mine.hi = -1;
  mine.wlo = mine.whi = 0.f;
  for (int k = ln; k < P * G; k += 32) {               // one iteration on the staged path (P * G <= 32)
    const int p = k / G, i = k - p * G;
    const float start = isx ? g.start_w : g.start_h, bin = isx ? g.bin_w : g.bin_h;
    const float v = start + p * bin + static_cast<float>(i + .5f) * bin / static_cast<float>(G);
    const AxisTap t = axis_tap(v, isx ? W : H);
    TapS o;
    o.lo = t.valid ? t.lo : -1;
    o.hi = t.valid ? t.hi : -1;
    o.wlo = t.wlo;
    o.whi = t.whi;
    out[(isx ? PH * G : 0) + k] = o;
    if (k == ln) mine = o;
  }
  if (clip != nullptr && P * G <= 32) {
    // the taps of the axis sit on lanes [0, P * G) and are monotone: per band, the first / last bin with a tap inside
    __syncwarp();
    const bool v = ln < P * G && mine.lo >= 0;
    const unsigned valid = __ballot_sync(0xffffffffu, v);
    if (valid) {
      const int lo0 = __shfl_sync(0xffffffffu, mine.lo, __ffs(valid) - 1);
      const int hi1 = __shfl_sync(0xffffffffu, mine.hi, 31 - __clz(valid));
      const int b0 = lo0 >> 3;
      for (int b = b0; b <= (hi1 >> 3) && b < NB; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, v && mine.hi >= 8 * b && mine.lo < 8 * b + 8);
        const int cv = m ? ((__ffs(m) - 1) / G) | (((31 - __clz(m)) / G) << 8) : 1;
        if (ln == 0) {
          if (m) clip[(long)r * NB + b] = cv;
          if (b - b0 < 64) s_clip[grp][isx][b - b0] = cv;
        }
      }
      if (ln == 0) {
        s_b0[grp][isx] = b0;
        s_nb[grp][isx] = min(min(hi1 >> 3, NB - 1) - b0 + 1, 64);
      }
    }
  }
  if (tile_count != nullptr) {
    // ---- the RoI appends itself to the candidate list of every 8x8 tile it has taps in: {RoI, clipped bin ranges}.
    //      The order inside a list is that of the atomics; the tile kernel sorts its (<= kCandCap) entries by RoI ----
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
    const int nby = s_nb[grp][0], nbx = s_nb[grp][1];
    const int tiles_x = tg8.tiles_x[l], per_img = tiles_x * tg8.tiles_y[l];
    const int tbase = tg8.first[pv.num_levels - 1 - l] + (int)roi[0] * per_img;
    for (int e = tid; e < nby * nbx; e += 64) {
      const int iy = e / nbx, ix = e - iy * nbx;
      const int rc = s_clip[grp][0][iy], cc = s_clip[grp][1][ix];
      if ((rc & 255) <= (rc >> 8) && (cc & 255) <= (cc >> 8)) {
        const int t = tbase + (s_b0[grp][0] + iy) * tiles_x + s_b0[grp][1] + ix;
        const int pos = atomicAdd(tile_count + t, 1);
        if (pos < kCandCap)
          cands[(long)t * kCandCap + pos] = make_int2(r, (rc & 255) | ((rc >> 8) << 8) | ((cc & 255) << 16) | ((cc >> 8) << 24));
      }
    }
  }
  if (tid == 0) {
    // sample coordinates are monotone along an axis, so the first / last sample bound the reached pixels
    const float yf = g.start_h + static_cast<float>(.5f) * g.bin_h / static_cast<float>(G);
    const float yl = g.start_h + (PH - 1) * g.bin_h + static_cast<float>(G - 1 + .5f) * g.bin_h / static_cast<float>(G);
    const float xf = g.start_w + static_cast<float>(.5f) * g.bin_w / static_cast<float>(G);
    const float xl = g.start_w + (PW - 1) * g.bin_w + static_cast<float>(G - 1 + .5f) * g.bin_w / static_cast<float>(G);
    const bool none = yl < -1.0f || yf > (float)H || xl < -1.0f || xf > (float)W;
    const int ylo = yf <= 0.f ? 0 : min((int)yf, H - 1), yhi = yl >= (float)(H - 1) ? H - 1 : (int)fmaxf(yl, 0.f) + 1;
    const int xlo = xf <= 0.f ? 0 : min((int)xf, W - 1), xhi = xl >= (float)(W - 1) ? W - 1 : (int)fmaxf(xl, 0.f) + 1;
    box[r] = none ? make_int4(1, 0, 1, 0) : make_int4(ylo, yhi, xlo, xhi);
    if (segid) segid[r] = l * pv.batch + (int)roi[0];
  }
}

// One launch prepares the whole call: blocks [0, L*B) bin the RoIs by (level, image); every later block builds the
// tap tables of 4 RoIs (64 threads each).
__global__ void __launch_bounds__(256) bwd_prepare(PyramidView pv, const float* __restrict__ rois, int K, int PH, int PW, int G,
                                                    int aligned, MapperView mp, const int* __restrict__ roi_levels,
                                                    int* __restrict__ seg_count, int* __restrict__ perm,
                                                    TapS* __restrict__ taps, int4* __restrict__ box, int* __restrict__ segid,
                                                    int* __restrict__ rowclip, int NBy, int* __restrict__ colclip, int NBx,
                                                    TileGrid tg8, int* __restrict__ tile_count, int2* __restrict__ cands) {
  const int nseg = pv.num_levels * pv.batch;
  if ((int)blockIdx.x < nseg) {
    bin_rois_block(pv, rois, K, mp, roi_levels, seg_count, perm, blockIdx.x);
  } else {
    const int r = 4 * ((int)blockIdx.x - nseg) + (threadIdx.x >> 6);
    if (r < K)
      roi_taps_group(pv, rois, PH, PW, G, aligned, mp, roi_levels, taps, box, r, threadIdx.x & 63, segid, rowclip, NBy, colclip,
                     NBx, tg8, tile_count, cands);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// deterministic tile-owner gather, staged: the single-pass kernel (no channel-vector copy of grad_out in HBM)
// ------------------------------------------------------------------------------------------------------------------
// One CTA per 8x8-pixel tile x 128-channel chunk, a warp per tile row, a lane per 4 channels (as bwd_tiles).  For every
// RoI that reaches the tile, in RoI-index order, the CTA
//   - clips the RoI's bins to those with a tap inside the tile (rows pA..pB, columns qA..qB),
//   - stages that sub-block of grad_out[r, c0:c0+128, :, :] in shared memory as channel-vector rows S[bin][128 ch] with
//     cp.async: 4-byte copies that transpose the reference's (K, C, PH, PW) layout on the fly (16-byte copies when the
//     pooled gradient is channels-last), double-buffered so the copies of the next sub-block fly under the arithmetic of
//     the current one (at most NBUF bins per buffer; larger sub-blocks are split by bin rows),
//   - builds the dense separable weight tables WY[tile row][bin row], WX[bin column][tile column]
//     (bilinear_interpolate_gradient :113-171 factors per axis), and
//   - accumulates  g[y][x][c] += sum_q WX[q][x] * (sum_p WY[y][p] * S[p][q][c])  in registers (packed FFMA2), vertical
//     combine first, in a fixed order: deterministic, no atomics.
namespace bst {
#ifndef BWD_TRACE
#define BWD_TRACE 0
#endif
#if BWD_TRACE
__device__ unsigned long long g_btrace[8192 * 8];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define BTR(i) do { if (threadIdx.x == 0 && blockIdx.x < 8192) g_btrace[blockIdx.x * 8 + (i)] = gtime(); } while (0)
#define BTRV(i, v) do { if (threadIdx.x == 0 && blockIdx.x < 8192) g_btrace[blockIdx.x * 8 + (i)] = (unsigned long long)(v); } while (0)
#else
#define BTR(i)
#define BTRV(i, v)
#endif

// Bins per staging buffer.  48 rather than 56: at 62 KB per CTA three CTAs fit the 196 KB shared-memory carve-out, which
// leaves the SM a 60 KB L1 instead of 28 KB -- the 4-byte transposing copies of a (K,C,PH,PW) gradient hit it (neighbouring
// bins of a channel share 32-byte sectors): 14x14 0.266 -> 0.223 ms, 7x7 0.131 -> 0.127 ms.  (A channels-last 14x14
// gradient, staged with 16-byte copies, would rather have 56 -- 0.190 vs 0.195 ms -- but both layouts keep the same item
// boundaries, so that their results stay bit-identical.)
constexpr int NBUF = 48;
constexpr int SROW = kChunk + 4;       // floats per staged bin (528 B: rows stay 16-byte aligned, 4-byte stores conflict-free)
constexpr int kRound = 2 * kTileThreads;   // RoIs of the (level, image) list scanned per round
constexpr int MAXQ = 32, MAXP = 16;    // bin columns / bin rows of one item the tables hold (the path needs P * G <= 32)

// per bin-column count nq (1..32): bin rows per item min(MAXP, NBUF / nq), and the multiplier with
// pq / nq == (pq * magic) >> 16 for pq < 64 (the item generator runs on one thread: no integer divisions there)
struct ItemTables {
  unsigned char npg[33];
  unsigned int magic[33];
  constexpr ItemTables() : npg(), magic() {
    for (int n = 1; n <= 32; n++) {
      npg[n] = (unsigned char)(NBUF / n < MAXP ? NBUF / n : MAXP);
      magic[n] = (65536u + n - 1) / n;
    }
  }
};
__constant__ ItemTables kItemTables = ItemTables();

struct __align__(16) Item {
  long long gofs;                      // float offset of the sub-block's first element inside grad_out
  int r, p0, np, q0, nq, nb;
  unsigned magic;                      // pq / nq == (pq * magic) >> 16 for pq < 64, nq <= 32
  int valid;
  int pad[2];
};

struct Smem {
  float S[2][NBUF * SROW];
  float2 WX[2][MAXQ][TW];              // weights duplicated for FFMA2
  float2 WY[2][TH][MAXP];
  int2 prange[2][TH];                  // first / last bin row of the item with weight on the tile row
  Item ring[4];
  int cand[kRound];
  int crange[kRound];                  // pA | pB << 8 | qA << 16 | qB << 24, or -1
  int wsum[8];
};

// WIDE: the L2 fetches 256 bytes around the word -- the rest of a 784-byte channel row of a 14x14 gradient is wanted by the
// next few copies anyway (0.272 -> 0.268 ms); for the 196-byte rows of a 7x7 gradient it only costs (0.131 -> 0.133 ms)
template <bool WIDE>
__device__ __forceinline__ void cp_async4(uint32_t sdst, const float* gsrc) {
  if (WIDE) asm volatile("cp.async.ca.shared.global.L2::256B [%0], [%1], 4;" ::"r"(sdst), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t sdst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// GO_CL: grad_out is (K, PH, PW, C) (a channels_last pooled gradient); otherwise (K, C, PH, PW) contiguous.
// PC / GC: compile-time pooled size (PH == PW == PC) and sampling grid, 0 = runtime values.
template <bool GO_CL, int PC, int GC>
__global__ void __launch_bounds__(kTileThreads, 3)
bwd_tiles_staged(PyramidView pv, TileGrid tg, const float* __restrict__ go, const TapS* __restrict__ taps,
                 const int4* __restrict__ box, int K, int PH_, int PW_, int G_, const int* __restrict__ seg_count,
                 const int* __restrict__ perm, int chunks, const int* __restrict__ tile_count,
                 const int2* __restrict__ cands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int PH = PC ? PC : PH_, PW = PC ? PC : PW_, G = GC ? GC : G_;
  const int C = pv.channels;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x / chunks;
  const int tile_id = t;
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
  const int cc = min(kChunk, C - c0);
  const bool active = 4 * lane < cc;
  const int PP = PH * PW;
  const int nt = (PH + PW) * G;
  const float invG = 1.0f / (float)G;
  const int y = y0 + warp;

  BTR(0);
  int n_items = 0;
  u64 acc[TW][2];
#pragma unroll
  for (int x = 0; x < TW; x++) acc[x][0] = acc[x][1] = 0ull;

  // the tile's candidates: appended by bwd_prepare (RoI + clipped bin ranges; one load + a rank sort), or -- a tile busier than the
  // prepared list holds, or a call without the lists -- found by scanning the (level, image) RoI list
  const int pc = tile_count ? tile_count[tile_id] : -1;
  const bool prepared = pc >= 0 && pc <= kCandCap;
  const int seg = l * pv.batch + b;
  const int nseg = prepared ? pc : seg_count[seg];
  const int* plist = perm + (long)seg * K;

  // staging constants of this thread
  //   transposing 4-byte copies: lane = (channel cl of a group of 4, bin j of a group of 8); the warp takes channel
  //   groups warp, warp + 8, warp + 16, warp + 24, i.e. channels cth + 32 k
  const int cl = lane & 3, j = lane >> 2;
  const int cth = 4 * warp + cl;
  const uint32_t s_base = smem_u32(&sm.S[0][0]);
  constexpr uint32_t kBufBytes = NBUF * SROW * 4;

  for (int base = 0; base < nseg; base += kRound) {
    int ncand = 0;
    if (prepared) {
      // the list was appended to with atomics: ranking the (distinct) RoI indices restores RoI order = the summation order
      int2 e = make_int2(0, 0);
      int* scratch = &sm.cand[kRound - kCandCap];              // the tail of the candidate array (ranks stay below it)
      if ((int)threadIdx.x < pc) {
        e = cands[(long)tile_id * kCandCap + threadIdx.x];
        scratch[threadIdx.x] = e.x;
      }
      __syncthreads();
      if ((int)threadIdx.x < pc) {
        int rank = 0;
        for (int i = 0; i < pc; i++) rank += scratch[i] < e.x;
        sm.cand[rank] = e.x;
        sm.crange[rank] = e.y;
      }
      ncand = pc;
      __syncthreads();
    } else {
    // ---- RoIs of this (level, image) whose reach intersects the tile, in RoI order (two list entries per thread) ----
    bool hit[2] = {false, false};
    int me[2] = {-1, -1};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = base + 2 * threadIdx.x + e;
      if (i < nseg) {
        me[e] = plist[i];
        const int4 bx = __ldg(box + me[e]);
        hit[e] = bx.x <= bx.y && bx.x < y0 + TH && bx.y >= y0 && bx.z < x0 + TW && bx.w >= x0;
      }
    }
    const unsigned bal0 = __ballot_sync(0xffffffffu, hit[0]), bal1 = __ballot_sync(0xffffffffu, hit[1]);
    if (lane == 0) sm.wsum[warp] = __popc(bal0) + __popc(bal1);
    __syncthreads();
    const unsigned below = (1u << lane) - 1;
    int pos = __popc(bal0 & below) + __popc(bal1 & below);
    for (int w = 0; w < 8; w++) {
      if (w < warp) pos += sm.wsum[w];
      ncand += sm.wsum[w];
    }
    if (hit[0]) sm.cand[pos] = me[0];
    if (hit[1]) sm.cand[pos + (hit[0] ? 1 : 0)] = me[1];
    __syncthreads();
    if (ncand == 0) continue;            // (uniform) nothing reaches the tile in this round
    // ---- per candidate: the bin rows / columns with a sample tap inside the tile (lane = sample) ----
    for (int ci = warp; ci < ncand; ci += 8) {
      const TapS* tp = taps + (long)sm.cand[ci] * nt;
      bool ty = false, tx = false;
      if (lane < PH * G) {
        const TapS s = ld_tap(tp + lane);
        ty = s.lo >= 0 && s.hi >= y0 && s.lo < y0 + TH;
      }
      if (lane < PW * G) {
        const TapS s = ld_tap(tp + PH * G + lane);
        tx = s.lo >= 0 && s.hi >= x0 && s.lo < x0 + TW;
      }
      const unsigned my = __ballot_sync(0xffffffffu, ty), mx = __ballot_sync(0xffffffffu, tx);
      if (lane == 0) {
        int cr = -1;
        if (my != 0u && mx != 0u) {
          const int pA = (__ffs(my) - 1) / G, pB = (31 - __clz(my)) / G;
          const int qA = (__ffs(mx) - 1) / G, qB = (31 - __clz(mx)) / G;
          cr = pA | (pB << 8) | (qA << 16) | (qB << 24);
        }
        sm.crange[ci] = cr;
      }
    }
    __syncthreads();
    }

    // ---- the sub-blocks ("items") of the round: thread 0 walks the candidates and publishes descriptors two items
    //      ahead in a 4-slot ring; everybody else only reads them ----
    int gci = 0, gp = 0;                 // generator state (thread 0)
    auto generate = [&](Item& out) {
      for (;;) {
        if (gci >= ncand) {
          out.valid = 0;
          return;
        }
        const int cr = sm.crange[gci];
        if (cr >= 0) {
          const int pA = cr & 255, pB = (cr >> 8) & 255, qA = (cr >> 16) & 255, qB = (cr >> 24) & 255;
          if (gp < pA) gp = pA;
          if (gp <= pB) {
            const int nq = qB - qA + 1;
            const int np = min((int)kItemTables.npg[nq], pB - gp + 1);
            const int r = sm.cand[gci];
            out.r = r;
            out.p0 = gp;
            out.np = np;
            out.q0 = qA;
            out.nq = nq;
            out.nb = np * nq;
            out.magic = kItemTables.magic[nq];
            out.valid = 1;
            out.gofs = GO_CL ? ((long long)r * PP + gp * PW + qA) * C + c0
                             : ((long long)r * C + c0) * PP + gp * PW + qA;
            gp += np;
            return;
          }
        }
        gci++;
        gp = 0;
      }
    };
    auto issue = [&](const Item& it, int bf) {
      const TapS* tp = taps + (long)it.r * nt;
      if (warp < 4) {   // WX[q][x]: thread = (bin column, pair of tile columns)
        const int qq = threadIdx.x >> 2, xa = x0 + 2 * (threadIdx.x & 3);
        float w0 = 0.f, w1 = 0.f;
        if (qq < it.nq) {
          const TapS* tq = tp + PH * G + (it.q0 + qq) * G;
#pragma unroll
          for (int i = 0; i < (GC ? GC : 1); i++)
            for (int ii = i; ii < G; ii += (GC ? GC : 1)) {
              const TapS s = ld_tap(tq + ii);
              w0 += tap_weight(s, xa);
              w1 += tap_weight(s, xa + 1);
            }
          w0 *= invG;
          w1 *= invG;
        }
        *reinterpret_cast<float4*>(&sm.WX[bf][qq][2 * (threadIdx.x & 3)]) = make_float4(w0, w0, w1, w1);
      } else {          // WY[y][p]: half-warp = tile row, lane = bin row of the item
        const int tt = threadIdx.x - 128;
        const int yy = tt >> 4, pp = tt & 15;
        float w = 0.f;
        if (pp < it.np) {
          const TapS* tq = tp + (it.p0 + pp) * G;
#pragma unroll
          for (int i = 0; i < (GC ? GC : 1); i++)
            for (int ii = i; ii < G; ii += (GC ? GC : 1)) w += tap_weight(ld_tap(tq + ii), y0 + yy);
          w *= invG;
        }
        sm.WY[bf][yy][pp] = make_float2(w, w);
        const unsigned m = (__ballot_sync(0xffffffffu, w != 0.f) >> (lane & 16)) & 0xffffu;
        if (pp == 0) sm.prange[bf][yy] = m ? make_int2(__ffs(m) - 1, 31 - __clz(m)) : make_int2(1, 0);
      }
      const uint32_t sb = s_base + bf * kBufBytes;
      const float* gsrc = go + it.gofs;
      const int nq = it.nq, nb = it.nb;
      const unsigned magic = it.magic;
      if (GO_CL) {
        if (active)
          for (int pq = warp; pq < nb; pq += 8) {
            const int pp = (pq * magic) >> 16, qq = pq - pp * nq;
            cp_async16(sb + (pq * SROW + 4 * lane) * 4, gsrc + (long)(pp * PW + qq) * C + 4 * lane);
          }
      } else if (cc == kChunk) {
        const float* gth = gsrc + cth * PP;
        const uint32_t sth = sb + cth * 4;
#pragma unroll 2
        for (int pq = j; pq < nb; pq += 8) {
          const int pp = (pq * magic) >> 16, qq = pq - pp * nq;
          const float* s = gth + (pp * PW + qq);
          const uint32_t d = sth + pq * (SROW * 4);
          cp_async4<(PC >= 12)>(d, s);
          cp_async4<(PC >= 12)>(d + 128, s + 32 * PP);
          cp_async4<(PC >= 12)>(d + 256, s + 64 * PP);
          cp_async4<(PC >= 12)>(d + 384, s + 96 * PP);
        }
      } else {
        for (int pq = j; pq < nb; pq += 8) {
          const int pp = (pq * magic) >> 16, qq = pq - pp * nq;
          const float* s = gsrc + (pp * PW + qq);
          const uint32_t d = sb + pq * (SROW * 4);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int c = cth + 32 * k;
            if (c < cc) cp_async4<false>(d + c * 4, s + (long)c * PP);
          }
        }
      }
      cp_async_commit();
    };
    auto compute = [&](const Item& it, int bf) {
      if (y >= H) return;
      const int2 pr = sm.prange[bf][warp];
      if (pr.x > pr.y) return;
      const int nq = it.nq;
      const u64* wyp = reinterpret_cast<const u64*>(&sm.WY[bf][warp][0]);
      const float* Sb = sm.S[bf] + 4 * lane + pr.x * nq * SROW;
      for (int qq = 0; qq < nq; qq++, Sb += SROW) {
        u64 v0 = 0ull, v1 = 0ull;
        const float* sp = Sb;
#pragma unroll 2
        for (int pp = pr.x; pp <= pr.y; pp++, sp += nq * SROW) {
          const u64 w2 = wyp[pp];
          const ulonglong2 f = *reinterpret_cast<const ulonglong2*>(sp);
          v0 = fma2(w2, f.x, v0);
          v1 = fma2(w2, f.y, v1);
        }
        const ulonglong2* wx = reinterpret_cast<const ulonglong2*>(&sm.WX[bf][qq][0]);
#pragma unroll
        for (int k = 0; k < TW / 2; k++) {
          const ulonglong2 w = wx[k];
          acc[2 * k][0] = fma2(w.x, v0, acc[2 * k][0]);
          acc[2 * k][1] = fma2(w.x, v1, acc[2 * k][1]);
          acc[2 * k + 1][0] = fma2(w.y, v0, acc[2 * k + 1][0]);
          acc[2 * k + 1][1] = fma2(w.y, v1, acc[2 * k + 1][1]);
        }
      }
    };

    if (base == 0) BTR(1);
    if (threadIdx.x == 0) {
      generate(sm.ring[0]);
      generate(sm.ring[1]);
    }
    __syncthreads();
    Item cur = sm.ring[0];
    int bf = 0;
    if (cur.valid) issue(cur, 0);
    for (int i = 0; cur.valid; i++) {
      cp_async_wait_all();
      __syncthreads();
      const Item nxt = sm.ring[(i + 1) & 3];
      if (threadIdx.x == 0) generate(sm.ring[(i + 2) & 3]);
      if (nxt.valid) issue(nxt, bf ^ 1);
      compute(cur, bf);
      cur = nxt;
      bf ^= 1;
      n_items++;
    }
    __syncthreads();
  }
  BTR(2);

  // ---- the tile's gradient: written exactly once ----
  if (active && y < H) {
    if (pv.layout == CPM_LAYOUT_NCHW) {
      // the reference's gradient layout (ROIAlign_cuda.cu:451-452): a lane owns 4 channel planes, in each 8 consecutive
      // columns of row y -- one 32-byte sector per (lane, channel)
      float* dst = (float*)pv.ptr[l] + (((long)b * C + c0 + 4 * lane) * H + y) * W + x0;
      const long cs = (long)H * W;
      const bool v4 = (W & 3) == 0 && x0 + TW <= W;
      const bool v8 = (W & 7) == 0 && x0 + TW <= W;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float v[TW];
#pragma unroll
        for (int x = 0; x < TW; x++) {
          float lo, hi;
          unpack2(acc[x][j >> 1], lo, hi);
          v[x] = (j & 1) ? hi : lo;
        }
        float* row = dst + j * cs;
        if (v8) {
          st_global_v8(row, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        } else if (v4) {
          reinterpret_cast<float4*>(row)[0] = make_float4(v[0], v[1], v[2], v[3]);
          reinterpret_cast<float4*>(row)[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int x = 0; x < TW; x++)
            if (x0 + x < W) row[x] = v[x];
        }
      }
    } else {
      ulonglong2* dst = reinterpret_cast<ulonglong2*>((float*)pv.ptr[l] + (((long)b * H + y) * W + x0) * C + c0) + lane;
      const long C4 = C >> 2;
#pragma unroll
      for (int x = 0; x < TW; x++)
        if (x0 + x < W) dst[x * C4] = make_ulonglong2(acc[x][0], acc[x][1]);
    }
  }
  BTR(3);
  BTRV(4, n_items);
  BTRV(5, pc);
}

#if BWD_TRACE
}  // namespace bst
}  // namespace cpm
extern "C" __attribute__((visibility("default"))) int cpm_debug_bwd_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, cpm::bst::g_btrace, sizeof(unsigned long long) * n);
}
namespace cpm {
namespace bst {
#endif

typedef void (*StagedFn)(PyramidView, TileGrid, const float*, const TapS*, const int4*, int, int, int, int, const int*,
                         const int*, int, const int*, const int2*);

template <bool GO_CL>
static StagedFn pick_staged(int PH, int PW, int G) {
  if (PH == 7 && PW == 7 && G == 2) return bwd_tiles_staged<GO_CL, 7, 2>;
  if (PH == 14 && PW == 14 && G == 2) return bwd_tiles_staged<GO_CL, 14, 2>;
  return bwd_tiles_staged<GO_CL, 0, 0>;
}

}  // namespace bst

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace cpm

using namespace cpm;

struct BwdWs {
  size_t seg_count, perm, taps, box, segid, rowclip, colclip, tile_count, cands, tile_off, lists, total;
  int NBy, NBx;          // bands of 8 pixel rows / columns of the largest level (clip table widths); 0: no clip tables
  long tiles8;           // 8x8 tiles of the pyramid
};

// the single-pass staged kernel takes every fixed-grid pooler whose samples per axis fit one ballot
static bool bwd_staged_ok(int PH, int PW, int G) { return G >= 1 && PH * G <= 32 && PW * G <= 32; }

// Which deterministic tile kernel takes a call (measured on the benchmark workload): the staged
// 8x8-tile kernel; CPM_BWD_IMPL=tma selects the TMA tile kernel for the two CPM poolers (measured slower; kept for A/B
// measurements).  Read once per process.
static int bwd_impl_env() {
  static const int v = [] {
    const char* e = getenv("CPM_BWD_IMPL");
    if (e != nullptr && strcmp(e, "tma") == 0) return 1;
    if (e != nullptr && strcmp(e, "staged") == 0) return 2;
    return 0;
  }();
  return v;
}

static bool tma_shape_ok(const cpm_pyramid_t* p, int PH, int PW, int G) {
  if (!(btma::shape_ok(PH, PW, G) && p->dtype == CPM_F32 && p->channels % btma::CH == 0)) return false;
  return bwd_impl_env() == 1;
}

// p == nullptr: the plain layout (no clip tables, no per-tile lists: tile CTAs scan); else with the prepared per-tile
// candidate lists of the staged kernel and, when `tma`, the stage lists of the TMA kernel
static BwdWs bwd_layout(int64_t K, int L, int B, int PH, int PW, int G, const cpm_pyramid_t* p, bool tma) {
  BwdWs w;
  const size_t segs = (size_t)(L > 0 ? L : 1) * (size_t)(B > 0 ? B : 1);
  const size_t k = (size_t)(K > 0 ? K : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  w.NBy = w.NBx = 0;
  w.tiles8 = 0;
  if (p != nullptr && bwd_staged_ok(PH, PW, G)) {
    for (int l = 0; l < p->num_levels; l++) {
      const int by = (p->height[l] + 7) / 8, bx = (p->width[l] + 7) / 8;
      w.NBy = by > w.NBy ? by : w.NBy;
      w.NBx = bx > w.NBx ? bx : w.NBx;
      w.tiles8 += (long)p->batch * by * bx;
    }
  }
  const long tma_tiles = (p != nullptr && tma) ? btma::num_tiles(p) : 0;
  w.seg_count = take(segs * sizeof(int));
  w.perm = take(segs * k * sizeof(int));
  w.taps = take(k * (size_t)(PH + PW) * (size_t)(G > 0 ? G : 1) * sizeof(TapS));
  w.box = take(k * sizeof(int4));
  w.segid = take(w.NBy ? k * sizeof(int) : 0);
  w.rowclip = take(k * (size_t)w.NBy * sizeof(int));
  w.colclip = take(k * (size_t)w.NBx * sizeof(int));
  w.tile_count = take((size_t)(w.tiles8 > tma_tiles ? w.tiles8 : tma_tiles) * sizeof(int));
  w.cands = take((size_t)w.tiles8 * kCandCap * sizeof(int2));
  w.tile_off = take((size_t)tma_tiles * sizeof(int));
  w.lists = take(tma_tiles > 0 ? btma::list_entries(p, K, PH) * sizeof(int2) : 0);
  w.total = off;
  return w;
}

extern "C" size_t cpm_roi_align_backward_workspace_bytes(int64_t K, int num_levels, int batch, int channels, int pooled_h,
                                                         int pooled_w, int sampling_ratio) {
  (void)channels;
  return bwd_layout(K, num_levels, batch, pooled_h, pooled_w, sampling_ratio, nullptr, false).total;
}

extern "C" size_t cpm_roi_align_backward_workspace_bytes_pyr(const cpm_pyramid_t* grad_feat, int64_t K, int pooled_h,
                                                             int pooled_w, int sampling_ratio) {
  if (grad_feat == nullptr || grad_feat->num_levels < 1 || grad_feat->num_levels > CPM_MAX_LEVELS) return 0;
  return bwd_layout(K, grad_feat->num_levels, grad_feat->batch, pooled_h, pooled_w, sampling_ratio, grad_feat,
                    tma_shape_ok(grad_feat, pooled_h, pooled_w, sampling_ratio)).total;
}

extern "C" int cpm_roi_align_backward(const cpm_pyramid_t* grad_feat, const void* d_grad_out, const void* d_rois, int64_t K,
                                      int pooled_h, int pooled_w, int sampling_ratio, int aligned, int interpolation,
                                      const cpm_level_mapper_t* mapper, const int32_t* d_roi_levels, int mode,
                                      void* d_workspace, size_t workspace_bytes, void* stream) {
  return cpm_roi_align_backward_ex(grad_feat, d_grad_out, d_rois, K, pooled_h, pooled_w, sampling_ratio, aligned, interpolation,
                                   mapper, d_roi_levels, mode, CPM_POOLED_KCHW, d_workspace, workspace_bytes, stream);
}

extern "C" int cpm_roi_align_backward_ex(const cpm_pyramid_t* grad_feat, const void* d_grad_out, const void* d_rois, int64_t K,
                                         int pooled_h, int pooled_w, int sampling_ratio, int aligned, int interpolation,
                                         const cpm_level_mapper_t* mapper, const int32_t* d_roi_levels, int mode,
                                         int pooled_layout, void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_pyramid(grad_feat, "grad_feat");
  if (rc != CPM_OK) return rc;
  CPM_CHECK_ARG(pooled_layout == CPM_POOLED_KCHW || pooled_layout == CPM_POOLED_KHWC, "unknown pooled layout %d", pooled_layout);
  if (pooled_layout == CPM_POOLED_KHWC &&
      !(mode == CPM_BWD_DETERMINISTIC && bwd_staged_ok(pooled_h, pooled_w, sampling_ratio) && grad_feat->channels % 4 == 0 &&
        ((uintptr_t)d_grad_out & 15) == 0)) {
    set_error("a channels-last pooled gradient (CPM_POOLED_KHWC) is read by the deterministic staged kernel only "
              "(pooled size * sampling_ratio <= 32, C %% 4 == 0, 16-byte aligned grad_out)");
    return CPM_ERR_UNSUPPORTED;
  }
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

  // what the fast kernels need of the gradient pyramid: fp32, bilinear, channel vectors of 4, 16-byte aligned levels
  bool fast_f32 = grad_feat->dtype == CPM_F32 && interpolation == CPM_INTERP_BILINEAR && C % 4 == 0;
  for (int l = 0; fast_f32 && l < L; l++) fast_f32 = ((uintptr_t)grad_feat->d_level[l] & 15) == 0;
  const bool nhwc_f32 = fast_f32 && grad_feat->layout == CPM_LAYOUT_NHWC;
  const int chunks = (C + kChunk - 1) / kChunk;

  if (mode == CPM_BWD_DETERMINISTIC) {
    // the TMA kernel: the two CPM poolers, (K,C,PH,PW) gradient, NHWC or NCHW gradient pyramid
    bool tma = pooled_layout == CPM_POOLED_KCHW && interpolation == CPM_INTERP_BILINEAR &&
               tma_shape_ok(grad_feat, pooled_h, pooled_w, sampling_ratio) && ((uintptr_t)d_grad_out & 15) == 0;
    for (int l = 0; tma && l < L; l++) tma = ((uintptr_t)grad_feat->d_level[l] & 15) == 0;
    // a workspace sized with the plain query (no pyramid shapes) has no room for the per-tile lists: tile CTAs then scan
    BwdWs w = bwd_layout(K, L, B, pooled_h, pooled_w, sampling_ratio, grad_feat, tma);
    if (tma && btma::list_entries(grad_feat, K, pooled_h) >= (1ull << 31)) tma = false;
    if (d_workspace == nullptr || workspace_bytes < w.total || (long)w.tiles8 * kCandCap >= (1L << 31)) {
      tma = false;
      w = bwd_layout(K, L, B, pooled_h, pooled_w, sampling_ratio, nullptr, false);
    }
    if (!tma && !(fast_f32 && bwd_staged_ok(pooled_h, pooled_w, sampling_ratio))) {
      set_error("deterministic backward needs an fp32 gradient pyramid (NHWC or NCHW, 16-byte aligned levels), bilinear "
                "interpolation, C %% 4 == 0 and pooled size * sampling_ratio <= 32 (sampling_ratio >= 1); use "
                "CPM_BWD_ATOMIC otherwise");
      return CPM_ERR_UNSUPPORTED;
    }
    if (d_workspace == nullptr || workspace_bytes < w.total) {
      set_error("workspace too small: %zu < %zu bytes", workspace_bytes, w.total);
      return CPM_ERR_WORKSPACE;
    }
    if ((rc = check_device_ptr(d_workspace, "workspace")) != CPM_OK) return rc;
    char* wsb = (char*)d_workspace;
    int* seg_count = (int*)(wsb + w.seg_count);
    int* perm = (int*)(wsb + w.perm);
    TapS* taps = (TapS*)(wsb + w.taps);
    int4* box = (int4*)(wsb + w.box);
    const int Kp = K > 0 ? (int)K : 1;
    if (K > 0) {
      TileGrid tg8;
      const long tiles8 = make_tile_grid(tg8, grad_feat, TH, TW);
      const bool lists = !tma && w.NBy > 0 && w.NBy <= 64 && w.NBx <= 64;
      if (lists) CPM_CHECK_CUDA(cudaMemsetAsync(wsb + w.tile_count, 0, (size_t)tiles8 * sizeof(int), st));
      bwd_prepare<<<(unsigned)(L * B + (K + 3) / 4), 256, 0, st>>>(
          pv, (const float*)d_rois, (int)K, pooled_h, pooled_w, sampling_ratio, aligned, mp, d_roi_levels, seg_count, perm, taps,
          box, w.NBy ? (int*)(wsb + w.segid) : nullptr, w.NBy ? (int*)(wsb + w.rowclip) : nullptr, w.NBy,
          w.NBx ? (int*)(wsb + w.colclip) : nullptr, w.NBx, tg8, lists ? (int*)(wsb + w.tile_count) : nullptr,
          lists ? (int2*)(wsb + w.cands) : nullptr);
      CPM_CHECK_LAUNCH();
    } else {
      CPM_CHECK_CUDA(cudaMemsetAsync(seg_count, 0, (size_t)L * B * sizeof(int), st));
    }
    if (tma) {
      rc = btma::launch(grad_feat, pv, (const float*)d_grad_out, (int)K, pooled_h, taps, box, (const int*)(wsb + w.rowclip),
                        w.NBy, seg_count, perm, (int*)(wsb + w.tile_count), (int*)(wsb + w.tile_off), (int2*)(wsb + w.lists), st);
      if (rc != CPM_ERR_UNSUPPORTED) return rc;
      // no tensor-map encoder in this driver: the generic staged kernel below
    }
    TileGrid tg;
    const long tiles = make_tile_grid(tg, grad_feat, TH, TW);
    CPM_CHECK_ARG(tiles * chunks < (1L << 31), "gradient pyramid too large for one launch");
    const bst::StagedFn fn = pooled_layout == CPM_POOLED_KHWC ? bst::pick_staged<true>(pooled_h, pooled_w, sampling_ratio)
                                                              : bst::pick_staged<false>(pooled_h, pooled_w, sampling_ratio);
    {
      // opt-in to the dynamic shared memory, once per (instantiation, device, host thread)
      static thread_local const void* done_fn[8];
      static thread_local int done_dev[8], ndone = 0;
      int dev, hit = 0;
      CPM_CHECK_CUDA(cudaGetDevice(&dev));
      for (int i = 0; i < ndone; i++) hit |= done_fn[i] == (const void*)fn && done_dev[i] == dev;
      if (!hit) {
        CPM_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(bst::Smem)));
        if (ndone < 8) {
          done_fn[ndone] = (const void*)fn;
          done_dev[ndone++] = dev;
        }
      }
    }
    const bool lists = w.NBy > 0 && w.NBy <= 64 && w.NBx <= 64 && K > 0;
    fn<<<(unsigned)(tiles * chunks), kTileThreads, sizeof(bst::Smem), st>>>(
        pv, tg, (const float*)d_grad_out, taps, box, Kp, pooled_h, pooled_w, sampling_ratio, seg_count, perm, chunks,
        lists ? (const int*)(wsb + w.tile_count) : nullptr, lists ? (const int2*)(wsb + w.cands) : nullptr);
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
