// RoIAlign forward, hot path for the CPM head's square poolers (7x7 and 14x14, sampling_ratio 1 or 2) on sm_100a.
//
// Replaces RoIAlignForward (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:178-256) + the per-level loop of Pooler.forward
// (pet/rcnn/utils/poolers.py:117-131).  The reference evaluates, per output element, gh*gw samples x 4 taps.  Bilinear
// sampling followed by the bin average is separable and linear, so for one RoI
//        out[c][ph][pw] = sum_y sum_x  Ay[ph][y] * Ax[pw][x] * F[y][x][c]
// where Ay / Ax hold, per bin, the summed tap weights of its samples on each feature row / column (a few non-zeros per
// bin; sample coordinates, validity and clamping are exactly bilinear_interpolate's, :36-86).  The kernel walks that
// product in the order that touches every feature pixel once per bin row:
//   CTA   = (RoI, channel chunk); the RoI's FPN level (LevelMapper, poolers.py:29-40) is computed in the kernel.
//   table = built once per CTA in shared memory: the distinct feature COLUMNS the RoI touches, in order, each with the
//           first bin it contributes to and its weights on that bin and the next NW-1 (NW = 2, 4 or 7, whichever covers
//           the RoI: small RoIs put many bins on one pixel).
//   warp  = one bin row (7x7) or two bin rows on half-warps (14x14) x one group of 7 bin columns; lane = 4 channels.
//           Per column: <= 4 coalesced LDG.128 (the distinct feature rows of the bin row's samples), a vertical
//           combine V = sum_k wy[k] * F[row_k][x] (packed FFMA2), then acc[pw + k] += w[k] * V with pw the STATIC
//           index of an unrolled loop over bins whose trip count is "columns first touched by bin pw" -- all
//           accumulators stay in registers, no dynamic register indexing, no per-sample control flow.
//   out   = finished bins go to a shared-memory tile laid out exactly like the (K, C, PH, PW) output; the tile leaves
//           with cp.async.bulk (one 25 KB bulk store per CTA at 7x7, one 784 B store per channel at 14x14).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace cpm {

typedef unsigned long long u64;

namespace fwdc {

constexpr int kBins = 7;         // bin columns per group
constexpr int kColMax = 64;      // distinct feature columns one RoI can touch: 2 taps x (<= 32 samples)

struct __align__(16) TapS {
  int lo, hi;       // lo < 0: sample out of range (contributes 0, ROIAlign_cuda.cu:46-49)
  float wlo, whi;
};

__device__ __forceinline__ TapS make_tap(float v, int size) {
  const AxisTap t = axis_tap(v, size);
  TapS r;
  r.lo = t.valid ? t.lo : -1;
  r.hi = t.valid ? t.hi : -1;
  r.wlo = t.wlo;
  r.whi = t.whi;
  return r;
}

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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared -> global bulk copy (UBLKCP); the issuing thread commits and later waits for the shared-memory reads
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Storage type of the pyramid and of the pooled output: fp32, or bf16 (4 channels of a lane = one 8-byte tap, widened to
// fp32 on arrival; all arithmetic is fp32, a finished bin is rounded to bf16 once).
template <bool BF> struct Store { typedef float T; typedef ulonglong2 Tap; };
template <> struct Store<true> { typedef __nv_bfloat16 T; typedef uint2 Tap; };

// 4 channels of one lane as two packed fp32 pairs
__device__ __forceinline__ void widen(const ulonglong2& t, u64& lo, u64& hi) {
  lo = t.x;
  hi = t.y;
}
__device__ __forceinline__ void widen(const uint2& t, u64& lo, u64& hi) {   // bf16 -> fp32: the bits move up by 16
  lo = ((u64)(t.x & 0xffff0000u) << 32) | (u64)(t.x << 16);
  hi = ((u64)(t.y & 0xffff0000u) << 32) | (u64)(t.y << 16);
}

template <int NG>
struct Tables {
  TapS xt[32];                      // x sample taps (NS = 7*NG*G <= 28)
  int xcol[kColMax];                // pixel column of table entry i (ascending)
  int xoff[kColMax];                // its byte offset inside a feature row: x * C * 4
  int bf[kColMax], bl[kColMax];     // first / last bin (0 .. 7*NG-1) with a tap on the column
  float2 w[NG][kColMax][kBins];     // per group: weight of column i on bins base+0 .. base+NW-1, duplicated for FFMA2
  int cnt[NG][8];                   // per group: number of columns whose first bin (inside the group) is pw
  int ib[NG], nw[NG];               // per group: first column index, weights per column (2, 4 or 7)
  int ncols;
};

// One warp's share: NR feature rows per lane, NW bins per column.
// OCL: the pooled output is channels-last (K, PH, PW, C): a finished bin is one 16-byte store per lane straight to global
// memory (tptr points at the lane's 4 channels of the group's first bin, ostride = C); otherwise it goes to the
// (channel, bin)-ordered shared-memory tile.
template <int NG, bool OCL>
__device__ __forceinline__ void store_bin(float* __restrict__ tptr, int pw, int ostride, u64 lo, u64 hi) {
  constexpr int PP = 49 * NG * NG;
  const float2 a = unpack2(lo), b = unpack2(hi);
  if (OCL) {
    *reinterpret_cast<float4*>(tptr + (size_t)pw * ostride) = make_float4(a.x, a.y, b.x, b.y);
  } else {
    tptr[pw] = a.x;
    tptr[PP + pw] = a.y;
    tptr[2 * PP + pw] = b.x;
    tptr[3 * PP + pw] = b.y;
  }
}
template <int NG, bool OCL>
__device__ __forceinline__ void store_bin(__nv_bfloat16* __restrict__ tptr, int pw, int ostride, u64 lo, u64 hi) {
  constexpr int PP = 49 * NG * NG;
  const float2 a = unpack2(lo), b = unpack2(hi);
  if (OCL) {
    const __nv_bfloat162 v0 = __floats2bfloat162_rn(a.x, a.y), v1 = __floats2bfloat162_rn(b.x, b.y);
    uint2 v;
    v.x = *reinterpret_cast<const unsigned*>(&v0);
    v.y = *reinterpret_cast<const unsigned*>(&v1);
    *reinterpret_cast<uint2*>(tptr + (size_t)pw * ostride) = v;
  } else {
    tptr[pw] = __float2bfloat16_rn(a.x);
    tptr[PP + pw] = __float2bfloat16_rn(a.y);
    tptr[2 * PP + pw] = __float2bfloat16_rn(b.x);
    tptr[3 * PP + pw] = __float2bfloat16_rn(b.y);
  }
}

template <int NG, int NR, int NW, bool OCL, bool BF>
__device__ __forceinline__ void run_columns(const char* __restrict__ base, const unsigned (&rowoff)[4], const u64 (&wy)[4],
                                            const Tables<NG>& tb, const int g, typename Store<BF>::T* __restrict__ tptr,
                                            const int ostride) {
  typedef typename Store<BF>::Tap Tap;
  constexpr int PP = 49 * NG * NG;
  u64 acc[kBins][2];
#pragma unroll
  for (int p = 0; p < kBins; p++) acc[p][0] = acc[p][1] = 0ull;
  int i = tb.ib[g];
#pragma unroll
  for (int pw = 0; pw < kBins; pw++) {
    int n = tb.cnt[g][pw];
#pragma unroll 1
    for (; n > 0; --n, ++i) {
      const unsigned xo = (unsigned)tb.xoff[i];
      Tap f[NR];
#pragma unroll
      for (int k = 0; k < NR; k++) f[k] = __ldg(reinterpret_cast<const Tap*>(base + (size_t)(rowoff[k] + xo)));
      u64 flo, fhi;
      widen(f[0], flo, fhi);
      u64 vlo = mul2(wy[0], flo), vhi = mul2(wy[0], fhi);
#pragma unroll
      for (int k = 1; k < NR; k++) {
        widen(f[k], flo, fhi);
        vlo = fma2(wy[k], flo, vlo);
        vhi = fma2(wy[k], fhi, vhi);
      }
      const u64* wp = reinterpret_cast<const u64*>(&tb.w[g][i][0]);
#pragma unroll
      for (int k = 0; k < NW; k++) {
        if (pw + k < kBins) {
          const u64 w = wp[k];
          acc[pw + k][0] = fma2(w, vlo, acc[pw + k][0]);
          acc[pw + k][1] = fma2(w, vhi, acc[pw + k][1]);
        }
      }
    }
    store_bin<NG, OCL>(tptr, pw, ostride, acc[pw][0], acc[pw][1]);
  }
}

// The same loop with the feature loads detached from the arithmetic: every lane streams its 16-byte taps through a private
// shared-memory ring with cp.async, kRingStages - 1 columns ahead (no register is tied up while a load is in flight and no
// lane ever reads another lane's slot, so cp.async.wait_group is the only synchronisation).
template <int NG>
struct RingCfg {
  static constexpr int kStages = NG == 1 ? 3 : 2;
  static constexpr int kRows = NG == 1 ? 4 : 3;             // feature rows per column a slot holds (NR above it: direct loads)
  static constexpr int kWarpBytes = kStages * kRows * 512;
};

__device__ __forceinline__ void cp_async16_ca(uint32_t sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
// L2-only variant: measured better for the 14x14 pooler (0.135 -> 0.129 ms: its rows are re-read by one other warp at
// most, and the L1 is the busiest unit), worse for 7x7 (0.070 -> 0.080 ms)
__device__ __forceinline__ void cp_async16_cg(uint32_t sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_ca(uint32_t sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NG, int NR, int NW, bool OCL, bool BF>
__device__ __forceinline__ void run_columns_ring(const char* __restrict__ base, const unsigned (&rowoff)[4], const u64 (&wy)[4],
                                                 const Tables<NG>& tb, const int g, typename Store<BF>::T* __restrict__ tptr,
                                                 const int ostride, ulonglong2* __restrict__ ring /* this lane's first slot */) {
  typedef typename Store<BF>::Tap Tap;     // (a bf16 tap uses the first 8 bytes of the lane's 16-byte slot)
  constexpr int PP = 49 * NG * NG;
  constexpr int S = RingCfg<NG>::kStages, NRM = RingCfg<NG>::kRows;
  const uint32_t ring_s = smem_u32(ring);
  u64 acc[kBins][2];
#pragma unroll
  for (int p = 0; p < kBins; p++) acc[p][0] = acc[p][1] = 0ull;
  int i = tb.ib[g];
  int tot = 0;
#pragma unroll
  for (int pw = 0; pw < kBins; pw++) tot += tb.cnt[g][pw];
  const int iend = i + tot;
  auto issue = [&](int col, int slot) {
    if (col < iend) {
      const unsigned xo = (unsigned)tb.xoff[col];
#pragma unroll
      for (int k = 0; k < NR; k++) {
        if (BF) cp_async8_ca(ring_s + (uint32_t)((slot * NRM + k) * 512), base + (size_t)(rowoff[k] + xo));
        else if (NG == 2) cp_async16_cg(ring_s + (uint32_t)((slot * NRM + k) * 512), base + (size_t)(rowoff[k] + xo));
        else cp_async16_ca(ring_s + (uint32_t)((slot * NRM + k) * 512), base + (size_t)(rowoff[k] + xo));
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < S - 1; st++) issue(i + st, st);
  int slot = 0, pre = S - 1;             // slot of column i, slot the next issue fills
#pragma unroll
  for (int pw = 0; pw < kBins; pw++) {
    int n = tb.cnt[g][pw];
#pragma unroll 1
    for (; n > 0; --n, ++i) {
      issue(i + S - 1, pre);
      cp_async_wait<S - 1>();
      Tap f[NR];
#pragma unroll
      for (int k = 0; k < NR; k++) f[k] = *reinterpret_cast<const Tap*>(ring + (slot * NRM + k) * 32);
      u64 flo, fhi;
      widen(f[0], flo, fhi);
      u64 vlo = mul2(wy[0], flo), vhi = mul2(wy[0], fhi);
#pragma unroll
      for (int k = 1; k < NR; k++) {
        widen(f[k], flo, fhi);
        vlo = fma2(wy[k], flo, vlo);
        vhi = fma2(wy[k], fhi, vhi);
      }
      const u64* wp = reinterpret_cast<const u64*>(&tb.w[g][i][0]);
#pragma unroll
      for (int k = 0; k < NW; k++) {
        if (pw + k < kBins) {
          const u64 w = wp[k];
          acc[pw + k][0] = fma2(w, vlo, acc[pw + k][0]);
          acc[pw + k][1] = fma2(w, vhi, acc[pw + k][1]);
        }
      }
      slot = slot + 1 == S ? 0 : slot + 1;
      pre = pre + 1 == S ? 0 : pre + 1;
    }
    store_bin<NG, OCL>(tptr, pw, ostride, acc[pw][0], acc[pw][1]);
  }
  cp_async_wait<0>();
}

template <int NG, int NR, bool OCL, bool BF>
__device__ __forceinline__ void run_nw_ring(const char* base, const unsigned (&rowoff)[4], const u64 (&wy)[4], const Tables<NG>& tb,
                                            const int g, typename Store<BF>::T* tptr, const int ostride, ulonglong2* ring) {
  const int nw = tb.nw[g];
  if (nw == 2) run_columns_ring<NG, NR, 2, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride, ring);
  else if (nw == 4) run_columns_ring<NG, NR, 4, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride, ring);
  else run_columns_ring<NG, NR, 7, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride, ring);
}

template <int NG, int NR, bool OCL, bool BF>
__device__ __forceinline__ void run_nw(const char* base, const unsigned (&rowoff)[4], const u64 (&wy)[4], const Tables<NG>& tb,
                                       const int g, typename Store<BF>::T* tptr, const int ostride) {
  const int nw = tb.nw[g];
  if (nw == 2) run_columns<NG, NR, 2, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride);
  else if (nw == 4) run_columns<NG, NR, 4, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride);
  else run_columns<NG, NR, 7, OCL, BF>(base, rowoff, wy, tb, g, tptr, ostride);
}

// NG = 1: 7x7 pooler, CTA = 7 warps (one bin row each) x 128 channels.
// NG = 2: 14x14 pooler, CTA = 14 warps (7 row pairs x 2 column groups) x 64 channels; half-warps own different bin rows.
template <int NG, bool OCL, bool BF>
__global__ void __launch_bounds__(224 * NG, NG == 1 ? 3 : 2)
roi_align_fwd_cols(PyramidView pv, const float* __restrict__ rois, int G, int aligned, MapperView mp,
                   const int* __restrict__ roi_levels, typename Store<BF>::T* __restrict__ out, int chunks, int cpc) {
  typedef typename Store<BF>::T T;
  constexpr int ES = (int)sizeof(T);       // bytes per stored element
  constexpr int P = kBins * NG;            // pooled height == width
  constexpr int PP = P * P;
  constexpr int LPR = 32 / NG;             // lanes per bin row
  constexpr int CH = 4 * LPR;              // channels per CTA
  constexpr int SK = NG == 1 ? 0 : 16 / ES;   // tile skew (16 bytes) per 4 channels: keeps the 14x14 tile's rows 16-byte aligned
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const tile = reinterpret_cast<T*>(smem_raw);     // [CH][PP] (+ skew), the CTA's block of the output
  __shared__ Tables<NG> tb;

  const int C = pv.channels;
  // One CTA pools `cpc` consecutive channel chunks of one RoI: the geometry, the column table and the bin-row taps are
  // built once and reused (they do not depend on the channel), only the column loop and the output store repeat.
  const int groups = chunks / cpc;
  const long n = blockIdx.x / groups;
  const int grp = blockIdx.x % groups;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* roi = rois + 5 * n;
  int l = 0;
  if (pv.num_levels > 1) l = roi_levels ? roi_levels[n] : fpn_level(roi[1], roi[2], roi[3], roi[4], mp);
  const RoiGeo<float> geo = roi_geometry<float>(roi, pv.scale[l < 0 || l >= pv.num_levels ? 0 : l], P, P, G, aligned != 0);
  const bool ok = l >= 0 && l < pv.num_levels && geo.b >= 0 && geo.b < pv.batch;
  constexpr int kTileElems = OCL ? 0 : CH * PP + SK * (CH / 4);       // a channels-last output needs no staging tile
  constexpr int kTileBytes = (kTileElems * ES + 15) & ~15;
  const bool store_issuer = NG == 1 ? threadIdx.x == 0 : threadIdx.x < CH / 4;  // threads that own bulk-store groups
  // tile -> out[n, c0 : c0 + CH, :, :], asynchronous: the next chunk's column loop runs while the tile drains
  auto store_tile = [&](int c0) {
    if (OCL) return;
    fence_async_smem();
    __syncthreads();
    T* o = out + ((size_t)n * C + c0) * PP;
    if (NG == 1) {
      if (threadIdx.x == 0) {
        bulk_s2g(o, tile, CH * PP * ES);
        bulk_commit();
      }
    } else {
      if (threadIdx.x < CH / 4) {      // the skew sits between groups of 4 channels: one 3 136-byte store per group
        const int c4 = threadIdx.x;
        bulk_s2g(o + (size_t)c4 * 4 * PP, tile + c4 * (4 * PP + SK), 4 * PP * ES);
        bulk_commit();
      }
    }
  };
  auto tile_free = [&](int ci) {        // before the tile is written again: its previous contents have left shared memory
    if (OCL || ci == 0) return;
    if (store_issuer) bulk_wait_read();
    __syncthreads();
  };
  if (!ok) {           // out-of-range level / image index: defined as zeros
    for (int ci = 0; ci < cpc; ci++) {
      const int c0 = (grp * cpc + ci) * CH;
      tile_free(ci);
      if (OCL) {
        for (int e = threadIdx.x; e < PP * (CH / 4); e += blockDim.x) {
          T* z = out + ((size_t)n * PP + e / (CH / 4)) * C + c0 + 4 * (e % (CH / 4));
          if (BF) *reinterpret_cast<uint2*>(z) = make_uint2(0u, 0u);
          else *reinterpret_cast<float4*>(z) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int e = threadIdx.x; e < kTileBytes / 4; e += blockDim.x) reinterpret_cast<unsigned*>(smem_raw)[e] = 0u;
      }
      store_tile(c0);
    }
  } else {
    const int H = pv.H[l], W = pv.W[l];
    const int NS = P * G;                  // samples per axis (<= 28)
    // ---- x table, step 1 (warp 0): sample taps -> ordered list of distinct columns, first / last bin per column ----
    if (warp == 0) {
      for (int e = lane; e < kColMax; e += 32) {
        tb.bf[e] = 1 << 20;
        tb.bl[e] = -1;
      }
      if (lane < NG * 8) (&tb.cnt[0][0])[lane] = 0;
      if (lane < NG) {
        tb.ib[lane] = 1 << 20;
        tb.nw[lane] = 0;          // holds the widest bin span of the group until it is turned into NW below
      }
      TapS t;
      t.lo = t.hi = -1;
      t.wlo = t.whi = 0.f;
      if (lane < NS) {
        const int p = lane / G, i = lane - p * G;
        const float v = geo.start_w + p * geo.bin_w + static_cast<float>(i + .5f) * geo.bin_w / static_cast<float>(G);
        t = make_tap(v, W);
      }
      tb.xt[lane] = t;
      const bool valid = t.lo >= 0;
      const int hi_prev = __shfl_up_sync(0xffffffffu, t.hi, 1);       // valid samples are one contiguous run
      const bool has_prev = lane > 0 && hi_prev >= 0;
      const bool newlo = valid && (!has_prev || t.lo > hi_prev);
      const bool newhi = valid && t.hi > t.lo && (!has_prev || t.hi > hi_prev);
      const unsigned blo = __ballot_sync(0xffffffffu, newlo), bhi = __ballot_sync(0xffffffffu, newhi);
      const unsigned le = 0xffffffffu >> (31 - lane);
      // the columns seen up to this sample end with ..., lo, hi (coordinates are monotone), so their table indices are
      const int cum = __popc(blo & le) + __popc(bhi & le);
      const int ihi = cum - 1, ilo = t.hi > t.lo ? cum - 2 : cum - 1;
      const int ncols = __popc(blo) + __popc(bhi);
      __syncwarp();
      if (newlo) {
        tb.xcol[ilo] = t.lo;
        tb.xoff[ilo] = t.lo * C * ES;
      }
      if (newhi) {
        tb.xcol[ihi] = t.hi;
        tb.xoff[ihi] = t.hi * C * ES;
      }
      if (valid) {
        const int bin = lane / G;
        atomicMin(&tb.bf[ilo], bin);
        atomicMax(&tb.bl[ilo], bin);
        if (t.hi > t.lo) {
          atomicMin(&tb.bf[ihi], bin);
          atomicMax(&tb.bl[ihi], bin);
        }
      }
      __syncwarp();
      // per group: column range, number of columns per first bin, widest bin span of a column
      for (int i = lane; i < ncols; i += 32) {
        const int bfi = tb.bf[i], bli = tb.bl[i];
#pragma unroll
        for (int g = 0; g < NG; g++) {
          if (bli >= kBins * g && bfi <= kBins * g + kBins - 1) {
            const int fb = max(bfi, kBins * g) - kBins * g;
            const int span = min(bli, kBins * g + kBins - 1) - kBins * g - fb;
            atomicAdd(&tb.cnt[g][fb], 1);
            atomicMax(&tb.nw[g], span);
            atomicMin(&tb.ib[g], i);
          }
        }
      }
      __syncwarp();
      if (lane < NG) {
        const int kmax = tb.nw[lane];
        const int nw = kmax <= 1 ? 2 : (kmax <= 3 ? 4 : 7);
        tb.nw[lane] = nw;
        if (nw == 7) {   // every column of the group carries its weight on all 7 bins, so all columns count for bin 0
          int tot = 0;
          for (int pw = 0; pw < kBins; pw++) {
            tot += tb.cnt[lane][pw];
            tb.cnt[lane][pw] = 0;
          }
          tb.cnt[lane][0] = tot;
        }
        if (tb.ib[lane] > ncols) tb.ib[lane] = ncols;
      }
      if (lane == 0) tb.ncols = ncols;
    } else if (warp == 1) {
      // ---- L2 prefetch of the RoI's footprint: one bulk prefetch per distinct feature row (all channels; the CTAs of the
      //      RoI's channel chunks share the rows round-robin).  It runs under the table build, so the main loop's loads
      //      find their lines in L2 instead of paying the DRAM latency once per column. ----
      const int NT = P * G;
      const float xf = geo.start_w + static_cast<float>(.5f) * geo.bin_w / static_cast<float>(G);
      const float xl = geo.start_w + (P - 1) * geo.bin_w + static_cast<float>(G - 1 + .5f) * geo.bin_w / static_cast<float>(G);
      const bool xany = !(xl < -1.0f || xf > (float)W);
      const int xmin = min(max((int)floorf(xf), 0), W - 1), xmax = min(max((int)floorf(xl) + 1, 0), W - 1);
      int lo = -1, hi = -1;
      if (lane < NT) {
        const int p = lane / G, i = lane - p * G;
        const float v = geo.start_h + p * geo.bin_h + static_cast<float>(i + .5f) * geo.bin_h / static_cast<float>(G);
        const AxisTap t = axis_tap(v, H);
        if (t.valid) {
          lo = t.lo;
          hi = t.hi;
        }
      }
      const int hi_prev = __shfl_up_sync(0xffffffffu, hi, 1);
      const bool has_prev = lane > 0 && hi_prev >= 0;
      const bool newlo = lo >= 0 && (!has_prev || lo > hi_prev);
      const bool newhi = lo >= 0 && hi > lo && (!has_prev || hi > hi_prev);
      const int chunk = grp;
      const int chunks = groups;       // the RoI's CTAs share the rows round-robin
      const char* img = reinterpret_cast<const char*>(pv.ptr[l]) + (size_t)geo.b * H * W * C * ES;
      const unsigned bytes = (unsigned)(xmax - xmin + 1) * (unsigned)C * (unsigned)ES;
      if (xany && newlo && lo % chunks == chunk) bulk_prefetch_l2(img + ((size_t)lo * W + xmin) * C * ES, bytes);
      if (xany && newhi && hi % chunks == chunk) bulk_prefetch_l2(img + ((size_t)hi * W + xmin) * C * ES, bytes);
    }
    // ---- this lane's bin row: distinct feature rows and their (count-normalised) weights ----
    const int g = NG == 1 ? 0 : warp / kBins;
    const int ph = NG == 1 ? warp : 2 * (warp % kBins) + (lane >> 4);
    const int lr = lane & (LPR - 1);
    unsigned rowoff[4] = {0u, 0u, 0u, 0u};
    float wyf[4] = {0.f, 0.f, 0.f, 0.f};
    int ry[4] = {-1, -1, -1, -1};
    int nrow = 0;
    {
      const float inv_count = 1.0f / (float)(G * G);
      for (int iy = 0; iy < G; iy++) {
        const float v = geo.start_h + ph * geo.bin_h + static_cast<float>(iy + .5f) * geo.bin_h / static_cast<float>(G);
        const AxisTap t = axis_tap(v, H);
        if (!t.valid) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int y = e ? t.hi : t.lo;
          const float wv = (e ? t.whi : t.wlo) * inv_count;
          if (e && t.hi == t.lo) continue;          // clamped at the last row: whi is 0 there
          bool found = false;
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (j < nrow && ry[j] == y) {
              wyf[j] += wv;
              found = true;
            }
          if (!found) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (j == nrow) {
                ry[j] = y;
                wyf[j] = wv;
              }
            nrow++;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int y = j < nrow ? ry[j] : (nrow > 0 ? ry[0] : 0);        // padding rows re-read a row already in L1, weight 0
        rowoff[j] = (unsigned)y * (unsigned)(W * C * ES);
        if (j >= nrow) wyf[j] = 0.f;
      }
    }
    u64 wy[4];
#pragma unroll
    for (int j = 0; j < 4; j++) wy[j] = pack2(wyf[j], wyf[j]);
    const int nr = max(2, __reduce_max_sync(0xffffffffu, nrow));
    __syncthreads();
    // ---- x table, step 2 (whole CTA): weights of every (group, column) on its NW bins ----
    {
      const int ncols = tb.ncols;
      for (int e = threadIdx.x; e < NG * ncols * kBins; e += blockDim.x) {
        const int k = e % kBins, i = (e / kBins) % ncols, gg = e / (kBins * ncols);
        const int nw = tb.nw[gg];
        const int b0 = kBins * gg + (nw == 7 ? 0 : max(tb.bf[i], kBins * gg) - kBins * gg);
        const int b = b0 + k;
        float w = 0.f;
        if (k < nw && b <= kBins * gg + kBins - 1 && b <= tb.bl[i] && b >= tb.bf[i]) {
          const int x = tb.xcol[i];
          for (int s = b * G; s < b * G + G; s++) {
            const TapS ts = tb.xt[s];
            if (ts.lo == x) w += ts.wlo;
            if (ts.hi == x && ts.hi != ts.lo) w += ts.whi;
          }
        }
        tb.w[gg][i][k] = make_float2(w, w);
      }
    }
    __syncthreads();
    // ---- main loop, once per channel chunk ----
    ulonglong2* ring = reinterpret_cast<ulonglong2*>(smem_raw + kTileBytes + warp * RingCfg<NG>::kWarpBytes) + lane;
    for (int ci = 0; ci < cpc; ci++) {
      const int c0 = (grp * cpc + ci) * CH;
      const char* base = reinterpret_cast<const char*>(pv.ptr[l]) + ((size_t)geo.b * H * W * C + c0 + 4 * lr) * ES;
      T* tptr = OCL ? out + ((size_t)n * PP + ph * P + kBins * g) * C + c0 + 4 * lr
                    : tile + (4 * lr) * PP + SK * lr + ph * P + kBins * g;
      tile_free(ci);
      if (nr == 2) run_nw_ring<NG, 2, OCL, BF>(base, rowoff, wy, tb, g, tptr, C, ring);
      else if (nr == 3) run_nw_ring<NG, 3, OCL, BF>(base, rowoff, wy, tb, g, tptr, C, ring);
      else if (RingCfg<NG>::kRows >= 4) run_nw_ring<NG, RingCfg<NG>::kRows >= 4 ? 4 : 2, OCL, BF>(base, rowoff, wy, tb, g, tptr, C, ring);
      else run_nw<NG, 4, OCL, BF>(base, rowoff, wy, tb, g, tptr, C);
      store_tile(c0);
    }
  }
  if (!OCL && store_issuer) bulk_wait_read();
}

}  // namespace fwdc

// true when the kernel above can take the call (checked by cpm_roi_align_forward)
bool fwd_cols_supported(const cpm_pyramid_t* feat, int pooled_h, int pooled_w, int sampling_ratio, const void* d_out) {
  if (feat->layout != CPM_LAYOUT_NHWC || (feat->dtype != CPM_F32 && feat->dtype != CPM_BF16)) return false;
  const double es = feat->dtype == CPM_BF16 ? 2.0 : 4.0;
  if (pooled_h != pooled_w || (pooled_h != 7 && pooled_h != 14)) return false;
  if (sampling_ratio != 1 && sampling_ratio != 2) return false;
  const int ch = pooled_h == 7 ? 128 : 64;
  if (feat->channels % ch != 0) return false;
  if (((uintptr_t)d_out & 15) != 0) return false;
  for (int l = 0; l < feat->num_levels; l++) {
    if (((uintptr_t)feat->d_level[l] & 15) != 0) return false;
    if ((double)feat->height[l] * feat->width[l] * feat->channels * es >= 2147483648.0) return false;   // 32-bit row offsets
  }
  return true;
}

// channel chunks pooled by one CTA (must divide the chunk count); CPM_FWD_CPC overrides it (A/B measurements; read once)
static int chunks_per_cta(int chunks, int want) {
  static const int forced = [] {
    const char* e = getenv("CPM_FWD_CPC");
    return e != nullptr && atoi(e) > 0 ? atoi(e) : 0;
  }();
  if (forced > 0) want = forced;
  while (want > 1 && chunks % want != 0) want--;
  return want < 1 ? 1 : want;
}

// opt-in to the dynamic shared memory of one instantiation, once per device and host thread
template <typename F>
static int configure_once(F fn, size_t smem, int& configured_dev) {
  int dev;
  CPM_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CPM_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev;
  }
  return CPM_OK;
}

template <bool BF>
static int launch_fwd_cols_t(const PyramidView& pv, const float* rois, long K, int P, int G, int aligned, const MapperView& mp,
                             const int* lv, void* out_, int out_channels_last, cudaStream_t st) {
  typedef typename fwdc::Store<BF>::T T;
  T* out = (T*)out_;
  const size_t ring1 = 7 * fwdc::RingCfg<1>::kWarpBytes, ring2 = 14 * fwdc::RingCfg<2>::kWarpBytes;
  const size_t smem1 = (out_channels_last ? 0 : (size_t)128 * 49 * sizeof(T)) + ring1;
  const size_t smem2 = (out_channels_last ? 0 : ((size_t)64 * 196 * sizeof(T) + 16 * 16)) + ring2;
  if (P == 7) {
    const int chunks = pv.channels / 128;
    CPM_CHECK_ARG(K * chunks < (1L << 31), "too many RoIs for one launch");
    auto fn = out_channels_last ? fwdc::roi_align_fwd_cols<1, true, BF> : fwdc::roi_align_fwd_cols<1, false, BF>;
    static thread_local int cfg[2] = {-1, -1};
    if (int rc = configure_once(fn, smem1, cfg[out_channels_last ? 1 : 0])) return rc;
    const int cpc = chunks_per_cta(chunks, 1);
    fn<<<(unsigned)(K * (chunks / cpc)), 224, smem1, st>>>(pv, rois, G, aligned, mp, lv, out, chunks, cpc);
  } else {
    const int chunks = pv.channels / 64;
    CPM_CHECK_ARG(K * chunks < (1L << 31), "too many RoIs for one launch");
    auto fn = out_channels_last ? fwdc::roi_align_fwd_cols<2, true, BF> : fwdc::roi_align_fwd_cols<2, false, BF>;
    static thread_local int cfg[2] = {-1, -1};
    if (int rc = configure_once(fn, smem2, cfg[out_channels_last ? 1 : 0])) return rc;
    const int cpc = chunks_per_cta(chunks, 2);
    fn<<<(unsigned)(K * (chunks / cpc)), 448, smem2, st>>>(pv, rois, G, aligned, mp, lv, out, chunks, cpc);
  }
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

// `out` has the pyramid's storage type (fp32 or bf16)
int launch_fwd_cols(const PyramidView& pv, const float* rois, long K, int P, int G, int aligned, const MapperView& mp,
                    const int* lv, void* out, int out_channels_last, cudaStream_t st) {
  return pv.dtype == CPM_BF16 ? launch_fwd_cols_t<true>(pv, rois, K, P, G, aligned, mp, lv, out, out_channels_last, st)
                              : launch_fwd_cols_t<false>(pv, rois, K, P, G, aligned, mp, lv, out, out_channels_last, st);
}

}  // namespace cpm
