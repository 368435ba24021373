// RoIAlign forward for the CPM head's square poolers (7x7 and 14x14, sampling_ratio 1 or 2) on sm_100a, row-streaming
// variant: every byte of a RoI's footprint crosses L2 -> SM once, carried by the TMA.
//
// Replaces RoIAlignForward (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:178-256) + the per-level loop of Pooler.forward
// (pet/rcnn/utils/poolers.py:117-131).  Bilinear sampling followed by the bin average is separable and linear:
//        out[c][p][q] = sum_y Ay[p][y] * ( sum_x Ax[q][x] * F[y][x][c] )
// (sample coordinates, validity and clamping exactly bilinear_interpolate's, :36-86, per axis).  roi_align_fwd_cols walks
// that product bin row by bin row and re-reads every feature row once per bin row that taps it (L2 -> SM traffic 4-5x the
// footprint).  Here the footprint streams through shared memory ONCE, row by row:
//   CTA      = (RoI, 1-4 channel chunks of 128 / 64 channels): 7 arithmetic warps + 1 producer warp.  The RoI's FPN level
//              (LevelMapper, poolers.py:29-40) is computed in the kernel and selects the level's tensor maps.  Every warp
//              derives the RoI's geometry for itself (no table exchange: one __syncthreads in the kernel, for the mbarrier
//              init); the tables are built once per RoI and the row stream runs on across its channel chunks.
//   producer = per PAIR of tapped feature rows (rows no sample touches are skipped), ONE cp.async.bulk.tensor.4d per row,
//              issued by lane 0 (UTMALDG; box {CH, 8n, 1, 1} of the NHWC map, one tensor map per level and segment width
//              8..48 pixels -- an instruction with per-lane operands would be replayed lane by lane), into a ring of slots,
//              completing on the slot's `full` mbarrier; a slot is refilled when the 7 arithmetic warps have arrived on its
//              `empty` mbarrier.  The producer starts fetching before the arithmetic warps have built their bin tables.
//   thread   = (bin column q, 4 channels).  Its x-taps (shared-memory offsets inside a row + weights) live in registers for
//              the whole RoI: 2G of them, or -- G = 2 and samples at most a pixel apart -- 3 pixels with combined weights.
//              Per row: those LDS.128 + packed FFMA2 -> T[y][q] in registers (the two rows of a slot are two independent
//              chains), kept in a 4-row register window (a bin row's samples tap at most 4 consecutive tapped rows); when
//              the last row of bin row p has passed, out[p][q] = sum_age wa[p][age] * window[age] goes to the shared-memory
//              output tile.  T never touches shared memory, nothing is exchanged between threads: no barrier in the loop.
//   out      = the (channel, bin)-ordered tile leaves with cp.async.bulk (UBLKCP) per chunk; the next chunk waits for its
//              shared-memory reads only in front of its first finished bin row.
// A RoI whose footprint does not fit the scheme (row segment wider than 48 pixels, more than 128 rows between its first
// and last tap, nothing valid) is pooled by the same CTA with the direct per-sample gather (reference arithmetic).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cpm {
namespace fwdr {

typedef unsigned long long u64;

constexpr int kBins = 7;
constexpr int kThreads = 256;       // 7 arithmetic warps + 1 producer warp
constexpr int kCompute = 224;
constexpr int kMaxChunks = 6;       // 8-pixel chunks of one row segment
constexpr int kMaxSlots = 16;
constexpr int kRowSpan = 128;       // tapped rows are tracked in a 128-bit mask from the first one

constexpr int kMaxMapLevels = 4;    // pyramids with more levels take roi_align_fwd_cols
// per level, one tensor map per width of the row segment (8, 16, ... 48 pixels): a row is ONE bulk tensor copy issued by
// one lane (an instruction with per-lane operands would be replayed lane by lane)
struct Maps {
  CUtensorMap m[kMaxMapLevels][kMaxChunks];
};

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
__device__ __forceinline__ ulonglong2 lds128(uint32_t a) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
// the same on shared-memory addresses (the loops keep them in registers)
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(b),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory"); }

template <int NG>
struct Cfg {
  static constexpr int P = kBins * NG, PP = P * P;
  static constexpr int CH = 128 / NG;             // channels per CTA
  static constexpr int QL = CH / 4;               // lanes per bin column
  static constexpr int SK = NG == 1 ? 0 : 4;      // tile skew (floats) per 4 channels: 2-way instead of 8-way store conflicts
  static constexpr int kTileBytes = ((CH * PP + SK * QL) * 4 + 127) & ~127;
  static constexpr int kChunkBytes = 8 * CH * 4;  // one TMA box: 8 pixels x CH channels
#ifndef FWDR_CTAS1
#define FWDR_CTAS1 2
#endif
#ifndef FWDR_RING1
#define FWDR_RING1 20
#endif
#ifndef FWDR_RING2
#define FWDR_RING2 28
#endif
  static constexpr int kRingBytes = (NG == 1 ? FWDR_RING1 : FWDR_RING2) * kChunkBytes;    // 80 KB / 56 KB: two CTAs per SM
  static constexpr int kSmemBytes = kTileBytes + kRingBytes;
  static_assert(kRingBytes >= 2 * kMaxChunks * kChunkBytes, "the ring must hold one pair of the widest rows");
};

// index of bit `i` among the set bits of the 128-bit mask m
__device__ __forceinline__ int rank128(const unsigned (&m)[4], int i) {
  int r = 0;
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const int lo = 32 * w;
    if (i >= lo + 32) r += __popc(m[w]);
    else if (i > lo) r += __popc(m[w] & ((1u << (i - lo)) - 1u));
  }
  return r;
}

#ifndef FWDR_TRACE
#define FWDR_TRACE 0
#endif
#if FWDR_TRACE
__device__ unsigned long long g_trace[8192 * 8];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TR(i) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 8192) g_trace[blockIdx.x * 8 + (i)] = gtime(); } while (0)
#else
#define TR(i)
#endif

// tile -> out[n, c0 : c0 + CH, :, :] by every thread of the CTA (the paths that do not stream), tile free on return
template <int NG>
__device__ __forceinline__ void store_tile_all(float* __restrict__ out, const float* tile, long n, int C, int c0) {
  typedef Cfg<NG> K_;
  const int tid = threadIdx.x;
  fence_async_smem();
  __syncthreads();
  float* o = out + ((size_t)n * C + c0) * K_::PP;
  if (NG == 1) {
    if (tid == 0) {
      bulk_s2g(o, tile, K_::CH * K_::PP * 4);
      bulk_commit();
      bulk_wait_read();
    }
  } else if (tid < K_::QL) {
    bulk_s2g(o + (size_t)tid * 4 * K_::PP, tile + tid * (4 * K_::PP + K_::SK), 4 * K_::PP * 4);
    bulk_commit();
    bulk_wait_read();
  }
  __syncthreads();
}

// Direct gather with the reference's arithmetic per output element (ROIAlign_cuda.cu:232-254), for the RoIs the streaming
// scheme does not take.  ty / tx: the taps of sample `lane` (every warp holds the same ones).
template <int NG, int G>
__device__ __noinline__ void pool_direct(const float* __restrict__ level, int H, int W, int C, int b, TapS ty, TapS tx,
                                         float* __restrict__ out, float* tile, TapS* yt, TapS* xt, long n, int cbase, int cpc) {
  typedef Cfg<NG> K_;
  constexpr int P = K_::P, PP = K_::PP, CH = K_::CH, QL = K_::QL, SK = K_::SK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) yt[lane] = ty;
  if (warp == 1) xt[lane] = tx;
  __syncthreads();
  for (int ci = 0; ci < cpc; ci++) {
    const int c0 = cbase + ci * CH;
    const float* img = level + (size_t)b * H * W * C + c0;
    for (int e = tid; e < PP * QL; e += kThreads) {
      const int bin = e / QL, quad = e - bin * QL;
      const int p = bin / P, q = bin - p * P;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int iy = 0; iy < G; iy++) {
        const TapS sy = yt[p * G + iy];
        if (sy.lo < 0) continue;
        for (int ix = 0; ix < G; ix++) {
          const TapS sx = xt[q * G + ix];
          if (sx.lo < 0) continue;
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(img + ((size_t)sy.lo * W + sx.lo) * C + 4 * quad));
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(img + ((size_t)sy.lo * W + sx.hi) * C + 4 * quad));
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(img + ((size_t)sy.hi * W + sx.lo) * C + 4 * quad));
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(img + ((size_t)sy.hi * W + sx.hi) * C + 4 * quad));
          const float w1 = sy.wlo * sx.wlo, w2 = sy.wlo * sx.whi, w3 = sy.whi * sx.wlo, w4 = sy.whi * sx.whi;
          acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
          acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
          acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
          acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
        }
      }
      const float cnt = (float)(G * G);
      float* t = tile + (4 * quad) * PP + SK * quad + bin;
      t[0] = acc.x / cnt;
      t[PP] = acc.y / cnt;
      t[2 * PP] = acc.z / cnt;
      t[3 * PP] = acc.w / cnt;
    }
    store_tile_all<NG>(out, tile, n, C, c0);
  }
}

template <int NG, int G>
__global__ void __launch_bounds__(kThreads, NG == 1 ? FWDR_CTAS1 : 2)
roi_align_fwd_rows(const __grid_constant__ Maps maps, PyramidView pv, const float* __restrict__ rois, int aligned, MapperView mp,
                   const int* __restrict__ roi_levels, float* __restrict__ out, int chunks, int cpc) {
  typedef Cfg<NG> K_;
  constexpr int P = K_::P, PP = K_::PP, CH = K_::CH, QL = K_::QL, SK = K_::SK;
  constexpr int NS = P * G;              // samples per axis (<= 28)
  constexpr int NTAP = 2 * G;            // taps per bin and axis
  constexpr float kInvG = 1.0f / (float)G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const tile = reinterpret_cast<float*>(smem_raw);          // [CH][PP] (+ skew), the CTA's block of the output
  __shared__ TapS xt[32], yt[32];        // direct-gather path only
  __shared__ int rows[kRowSpan];         // producer warp: tapped feature rows, ascending
  __shared__ __align__(16) float4 wa[8][16];  // per warp (private copy), per bin row: weight of the window row of age 0..3
  __shared__ __align__(8) uint64_t full[kMaxSlots], empty[kMaxSlots];

  const int C = pv.channels;
  // One CTA pools `cpc` consecutive channel chunks of one RoI: the tables are built once, the row stream runs on across
  // the chunk boundary (the producer is already fetching chunk i + 1 while chunk i's tile drains).
  const int groups = chunks / cpc;
  const long n = blockIdx.x / groups;
  const int cbase = (int)(blockIdx.x - n * groups) * cpc * CH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TR(0);
  const float* roi = rois + 5 * n;
  float rv[5];
#pragma unroll
  for (int i = 0; i < 5; i++) rv[i] = __ldg(roi + i);
  int l = 0;
  if (pv.num_levels > 1 && roi_levels) l = __ldg(roi_levels + n);
  if (tid == 0) {
    for (int s = 0; s < kMaxSlots; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kCompute / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();       // the only block-wide barrier of the streaming path (the RoI's loads are in flight across it)
  if (pv.num_levels > 1 && !roi_levels) l = fpn_level(rv[1], rv[2], rv[3], rv[4], mp);
  const RoiGeo<float> geo = roi_geometry<float>(rv, pv.scale[l < 0 || l >= pv.num_levels ? 0 : l], P, P, G, aligned != 0);
  const bool ok = l >= 0 && l < pv.num_levels && geo.b >= 0 && geo.b < pv.batch;
  if (!ok) {           // out-of-range level / image index: defined as zeros
    for (int e = tid; e < K_::kTileBytes / 4; e += kThreads) reinterpret_cast<unsigned*>(smem_raw)[e] = 0u;
    for (int ci = 0; ci < cpc; ci++) store_tile_all<NG>(out, tile, n, C, cbase + ci * CH);
    return;
  }
  const int H = pv.H[l], W = pv.W[l];

  // ---- every warp works out the whole geometry for itself: sample taps of both axes (sample = lane), the footprint's
  //      extent, the tapped rows as a 128-bit mask from the first one ----
  TapS ty, tx;
  ty.lo = ty.hi = tx.lo = tx.hi = -1;
  ty.wlo = ty.whi = tx.wlo = tx.whi = 0.f;
  if (lane < NS) {
    const int p = lane / G, i = lane - p * G;
    ty = make_tap(geo.start_h + p * geo.bin_h + static_cast<float>(i + .5f) * geo.bin_h / static_cast<float>(G), H);
    tx = make_tap(geo.start_w + p * geo.bin_w + static_cast<float>(i + .5f) * geo.bin_w / static_cast<float>(G), W);
  }
  const int ymin = __reduce_min_sync(0xffffffffu, ty.lo >= 0 ? ty.lo : 0x7fffffff);
  const int ymax = __reduce_max_sync(0xffffffffu, ty.hi);
  const int xlo = __reduce_min_sync(0xffffffffu, tx.lo >= 0 ? tx.lo : 0x7fffffff);
  const int xhi = __reduce_max_sync(0xffffffffu, tx.hi);
  const bool fits = ymax >= 0 && ymax - ymin < kRowSpan;
  unsigned m[4] = {0u, 0u, 0u, 0u};
  if (ty.lo >= 0 && fits) {
    const int a = ty.lo - ymin, b = ty.hi - ymin;
#pragma unroll
    for (int w = 0; w < 4; w++) {
      if ((a >> 5) == w) m[w] |= 1u << (a & 31);
      if ((b >> 5) == w) m[w] |= 1u << (b & 31);
    }
  }
#pragma unroll
  for (int w = 0; w < 4; w++) m[w] = __reduce_or_sync(0xffffffffu, m[w]);
  const int nr = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
  const int nch = (xhi - xlo + 8) >> 3;

  // (sample rows must not decrease with the bin row: a RoI of negative height, possible with aligned = true, does not stream)
  const bool slow = !fits || xhi < 0 || nch > kMaxChunks || nr == 0 || !(geo.bin_h >= 0.f);
  if (slow) {
    pool_direct<NG, G>(reinterpret_cast<const float*>(pv.ptr[l]), H, W, C, geo.b, ty, tx, out, tile, yt, xt, n, cbase, cpc);
    return;
  }
  TR(1);
  // a ring slot holds a PAIR of consecutive tapped rows (one `full` / `empty` round per pair; the two rows' arithmetic
  // chains are independent, which is what hides the shared-memory latency inside a warp)
  const int rowbytes = nch * K_::kChunkBytes;
  const int slotbytes = 2 * rowbytes;
  const int nslots = min(kMaxSlots, K_::kRingBytes / slotbytes);
  // (opaque to the compiler: it would otherwise rebuild these shared-window addresses from %cluster_ctaid at every use)
  uint32_t ring_s = smem_u32(smem_raw + K_::kTileBytes), full_s = smem_u32(full), empty_s = smem_u32(empty);
  asm volatile("" : "+r"(ring_s), "+r"(full_s), "+r"(empty_s));
  const uint32_t bar_end = (uint32_t)(nslots * 8);

  if (warp == 7) {
    // ---- producer: the row list, then one row segment per iteration, across the CTA's chunks ----
#pragma unroll
    for (int w = 0; w < 4; w++) {
      const int i = 32 * w + lane;
      if ((m[w] >> lane) & 1u) rows[rank128(m, i)] = ymin + i;
    }
    __syncwarp();
    const CUtensorMap* tm = &maps.m[l][nch - 1];
    uint32_t dst = ring_s, bo = 0;
    int wrap = 0;
    for (int ci = 0; ci < cpc; ci++) {
      const int c0 = cbase + ci * CH;
      for (int r = 0; r < nr; r += 2) {
        const int y0 = rows[r];
        const bool two = r + 1 < nr;
        const int y1 = two ? rows[r + 1] : 0;
        if (wrap > 0) mbar_wait_s(empty_s + bo, (unsigned)(wrap - 1) & 1u);
        if (lane == 0) {
          mbar_expect_tx_s(full_s + bo, (uint32_t)(two ? slotbytes : rowbytes));
          tma_load_4d(dst, tm, c0, xlo, y0, geo.b, full_s + bo);
          if (two) tma_load_4d(dst + (uint32_t)rowbytes, tm, c0, xlo, y1, geo.b, full_s + bo);
        }
        dst += (uint32_t)slotbytes;
        bo += 8;
        if (bo == bar_end) {
          bo = 0;
          dst -= (uint32_t)(nslots * slotbytes);
          wrap++;
        }
      }
    }
    TR(6);
    return;
  }

  // ---- arithmetic warps (the producer is already fetching) ----
  // per bin row (lane = bin row): the row that completes it and the window weights (ages 0..3: a bin row's samples tap at
  // most 4 consecutive entries of the row list)
  float a4[4] = {0.f, 0.f, 0.f, 0.f};
  int rl = -1;
  {
    int idx[NTAP];
    float w[NTAP];
#pragma unroll
    for (int i = 0; i < G; i++) {
      const int src = (lane * G + i) & 31;
      const int lo = __shfl_sync(0xffffffffu, ty.lo, src), hi = __shfl_sync(0xffffffffu, ty.hi, src);
      const float wl = __shfl_sync(0xffffffffu, ty.wlo, src), wh = __shfl_sync(0xffffffffu, ty.whi, src);
      const bool valid = lane < P && lo >= 0;
      idx[2 * i] = valid ? rank128(m, lo - ymin) : -1;
      idx[2 * i + 1] = valid ? rank128(m, hi - ymin) : -1;
      w[2 * i] = valid ? wl * kInvG : 0.f;
      w[2 * i + 1] = valid ? wh * kInvG : 0.f;
      rl = max(rl, idx[2 * i + 1]);
    }
    // a bin row without a valid sample is emitted (as zeros) with its predecessor
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, rl, d);
      if (lane >= d) rl = max(rl, o);
    }
    rl = max(rl, 0);
#pragma unroll
    for (int k = 0; k < NTAP; k++) {
      const int age = rl - idx[k];
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (idx[k] >= 0 && age == j) a4[j] += w[k];
    }
  }
  const int q = NG == 1 ? warp : 2 * warp + (lane >> 4);
  const int quad = lane & (QL - 1);
  uint32_t xo[NTAP];
  u64 xw[NTAP];
  // G = 2: when the four taps of every bin column fall on at most 3 consecutive pixels (sample spacing <= 1 pixel: bins up to
  // 2 pixels wide), a thread reads 3 pixels per row with combined weights instead of 4 (a quarter less shared-memory traffic)
  bool three = false;
  if (G == 2) {
    const int lo_o = __shfl_xor_sync(0xffffffffu, tx.lo, 1), hi_o = __shfl_xor_sync(0xffffffffu, tx.hi, 1);
    const int cmin = min(tx.lo >= 0 ? tx.lo : 0x7fffffff, lo_o >= 0 ? lo_o : 0x7fffffff), cmax = max(tx.hi, hi_o);
    three = __all_sync(0xffffffffu, cmax < 0 || cmax - cmin <= 2);
  }
  if (G == 2 && three) {
    int col[4];
    float wv[4];
#pragma unroll
    for (int i = 0; i < G; i++) {
      const int src = q * G + i;
      col[2 * i] = __shfl_sync(0xffffffffu, tx.lo, src);
      col[2 * i + 1] = __shfl_sync(0xffffffffu, tx.hi, src);
      const float wl = __shfl_sync(0xffffffffu, tx.wlo, src), wh = __shfl_sync(0xffffffffu, tx.whi, src);
      const bool valid = col[2 * i] >= 0;
      wv[2 * i] = valid ? wl * kInvG : 0.f;
      wv[2 * i + 1] = valid ? wh * kInvG : 0.f;
    }
    int cmin = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < 4; e += 2)
      if (col[e] >= 0) cmin = min(cmin, col[e]);
    if (cmin == 0x7fffffff) cmin = xlo;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float w = 0.f;
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (col[e] >= 0 && col[e] == cmin + k) w += wv[e];
      xo[k] = (uint32_t)((min(cmin + k, xhi) - xlo) * CH * 4 + quad * 16);
      xw[k] = pack2(w, w);
    }
    xo[3] = xo[0];
    xw[3] = 0ull;
  } else {
#pragma unroll
    for (int i = 0; i < G; i++) {
      const int src = q * G + i;
      const int lo = __shfl_sync(0xffffffffu, tx.lo, src), hi = __shfl_sync(0xffffffffu, tx.hi, src);
      const float wl = __shfl_sync(0xffffffffu, tx.wlo, src), wh = __shfl_sync(0xffffffffu, tx.whi, src);
      const bool valid = lo >= 0;
      xo[2 * i] = (uint32_t)((valid ? (lo - xlo) * CH * 4 : 0) + quad * 16);
      xo[2 * i + 1] = (uint32_t)((valid ? (hi - xlo) * CH * 4 : 0) + quad * 16);
      const float a = valid ? wl * kInvG : 0.f, b = valid ? wh * kInvG : 0.f;
      xw[2 * i] = pack2(a, a);
      xw[2 * i + 1] = pack2(b, b);
    }
  }
  // the completing rows of the 14 (7) bin rows, one byte each, in two registers; the window weights in the warp's table
  u64 rl_lo = 0ull, rl_hi = 0ull;
#pragma unroll
  for (int p = 0; p < P; p++) {
    const u64 v = (u64)(unsigned)__shfl_sync(0xffffffffu, rl, p);
    if (p < 8) rl_lo |= v << (8 * p);
    else rl_hi |= v << (8 * (p - 8));
  }
  if (lane < P) wa[warp][lane] = make_float4(a4[0], a4[1], a4[2], a4[3]);
  __syncwarp();
  const float4* const wap = wa[warp];
  float* const tptr = tile + (4 * quad) * PP + SK * quad + q;
  const bool issuer = NG == 1 ? tid == 0 : tid < QL;
  uint32_t base = ring_s, bo = 0;
  unsigned phase = 0;
  for (int ci = 0; ci < cpc; ci++) {
    const int c0 = cbase + ci * CH;
    u64 win[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++) win[j][0] = win[j][1] = 0ull;
    int pn = 0;
    u64 rlp = rl_lo;
    int next_last = (int)(rlp & 0xffull);
    // the previous chunk's tile must have left shared memory before the first bin row of this one is stored -- not before
    // its first feature rows are read: the wait sits in front of the first emission
    bool tile_busy = ci > 0;
    auto emit = [&](const u64 (&w0)[2], const u64 (&w1)[2], const u64 (&w2)[2], const u64 (&w3)[2]) {
      // bin row pn from the window rows of age 0 (the row just finished) .. 3
      if (tile_busy) {
        if (issuer) bulk_wait_read();
        bar_compute();
        tile_busy = false;
      }
      const float4 a = wap[pn];
      const u64 a0 = pack2(a.x, a.x), a1 = pack2(a.y, a.y), a2 = pack2(a.z, a.z), a3 = pack2(a.w, a.w);
      u64 ol = mul2(a0, w0[0]), oh = mul2(a0, w0[1]);
      ol = fma2(a1, w1[0], ol);
      oh = fma2(a1, w1[1], oh);
      ol = fma2(a2, w2[0], ol);
      oh = fma2(a2, w2[1], oh);
      ol = fma2(a3, w3[0], ol);
      oh = fma2(a3, w3[1], oh);
      const float2 v0 = unpack2(ol), v1 = unpack2(oh);
      float* t = tptr + pn * P;
      t[0] = v0.x;
      t[PP] = v0.y;
      t[2 * PP] = v1.x;
      t[3 * PP] = v1.y;
      pn++;
      rlp = pn == 8 ? rl_hi : rlp >> 8;
      next_last = pn < P ? (int)(rlp & 0xffull) : -1;
    };
    for (int r0 = 0; r0 < nr; r0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const int r = r0 + u;
        if (r < nr) {
          const bool two = r + 1 < nr;
          mbar_wait_s(full_s + bo, phase);
          if (r == 0 && ci == 0) TR(2);
          // an odd last row is read twice; its second copy lands in a window slot no emission reads any more
          const uint32_t base2 = base + (two ? (uint32_t)rowbytes : 0u);
          ulonglong2 f[NTAP], g[NTAP];
#pragma unroll
          for (int k = 0; k < NTAP; k++)
            if (k < 3 || !three) f[k] = lds128(base + xo[k]);
#pragma unroll
          for (int k = 0; k < NTAP; k++)
            if (k < 3 || !three) g[k] = lds128(base2 + xo[k]);
          u64 tl = mul2(xw[0], f[0].x), th = mul2(xw[0], f[0].y);
          u64 sl = mul2(xw[0], g[0].x), sh = mul2(xw[0], g[0].y);
#pragma unroll
          for (int k = 1; k < NTAP; k++) {
            if (k < 3 || !three) {
              tl = fma2(xw[k], f[k].x, tl);
              th = fma2(xw[k], f[k].y, th);
              sl = fma2(xw[k], g[k].x, sl);
              sh = fma2(xw[k], g[k].y, sh);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_s(empty_s + bo);
          base += (uint32_t)slotbytes;
          bo += 8;
          if (bo == bar_end) {
            bo = 0;
            base = ring_s;
            phase ^= 1u;
          }
          win[u][0] = tl;
          win[u][1] = th;
          while (next_last == r) emit(win[u], win[(u + 3) & 3], win[(u + 2) & 3], win[(u + 1) & 3]);     // uniform over the CTA
          win[u + 1][0] = sl;
          win[u + 1][1] = sh;
          while (next_last == r + 1) emit(win[u + 1], win[u], win[(u + 3) & 3], win[(u + 2) & 3]);
        }
      }
    }
    // ---- tile -> out[n, c0 : c0 + CH, :, :] ----
    if (ci == 0) TR(3);
    fence_async_smem();
    bar_compute();
    float* o = out + ((size_t)n * C + c0) * PP;
    if (NG == 1) {
      if (tid == 0) {
        bulk_s2g(o, tile, CH * PP * 4);
        bulk_commit();
      }
    } else if (tid < QL) {
      bulk_s2g(o + (size_t)tid * 4 * PP, tile + tid * (4 * PP + SK), 4 * PP * 4);
      bulk_commit();
    }
  }
  if (issuer) bulk_wait_read();
  TR(5);
#if FWDR_TRACE
  if (tid == 0 && blockIdx.x < 8192) g_trace[blockIdx.x * 8 + 7] = ((unsigned long long)nr << 32) | (unsigned)(nch | (nslots << 8));
#endif
}

#if FWDR_TRACE
}  // namespace fwdr
}  // namespace cpm
extern "C" __attribute__((visibility("default"))) int cpm_debug_fwd_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, cpm::fwdr::g_trace, sizeof(unsigned long long) * n);
}
namespace cpm {
namespace fwdr {
#endif

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor maps are a function of (pointer, shape, box): the 24 maps of a pyramid (4 levels x 6 segment widths) are encoded once
// per host thread and reused for as long as the same maps are pooled (the 5 poolers of a CPM iteration, steps whose tensors
// recur at the same addresses, every replay of a captured graph's launch).
struct PyramidKey {
  const void* ptr[kMaxMapLevels];
  int H[kMaxMapLevels], W[kMaxMapLevels];
  int L, B, C, CH;
};
struct PyramidMaps {
  PyramidKey key;
  Maps maps;
  bool used;
};

static int encode_map(CUtensorMap* dst, const void* ptr, int C, int W, int H, int B, int CH, int px) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CPM_ERR_UNSUPPORTED;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)CH, (cuuint32_t)px, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = enc(dst, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return CPM_ERR_UNSUPPORTED;
  }
  return CPM_OK;
}

static int pyramid_maps(const Maps** out, const cpm_pyramid_t* feat, int CH) {
  static thread_local PyramidMaps cache[8];
  static thread_local int next = 0;
  PyramidKey key;
  memset(&key, 0, sizeof(key));
  key.L = feat->num_levels;
  key.B = feat->batch;
  key.C = feat->channels;
  key.CH = CH;
  for (int l = 0; l < feat->num_levels; l++) {
    key.ptr[l] = feat->d_level[l];
    key.H[l] = feat->height[l];
    key.W[l] = feat->width[l];
  }
  for (int i = 0; i < 8; i++)
    if (cache[i].used && memcmp(&cache[i].key, &key, sizeof(key)) == 0) {
      *out = &cache[i].maps;
      return CPM_OK;
    }
  PyramidMaps& s = cache[next];
  s.used = false;
  memset(&s.maps, 0, sizeof(s.maps));
  for (int l = 0; l < feat->num_levels; l++)
    for (int k = 0; k < kMaxChunks; k++)
      if (int rc = encode_map(&s.maps.m[l][k], feat->d_level[l], feat->channels, feat->width[l], feat->height[l], feat->batch, CH,
                              8 * (k + 1)))
        return rc;
  s.key = key;
  s.used = true;
  next = (next + 1) & 7;
  *out = &s.maps;
  return CPM_OK;
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

template <int NG, int G>
static int launch_t(const Maps& maps, const PyramidView& pv, const float* rois, long K, int aligned, const MapperView& mp,
                    const int* lv, float* out, cudaStream_t st) {
  typedef Cfg<NG> K_;
  auto fn = roi_align_fwd_rows<NG, G>;
  static thread_local int configured_dev = -1;
  int dev;
  CPM_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CPM_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, K_::kSmemBytes));
    configured_dev = dev;
  }
  const int chunks = pv.channels / K_::CH;
  CPM_CHECK_ARG(K * chunks < (1L << 31), "too many RoIs for one launch");
  const int cpc = chunks_per_cta(chunks, NG == 1 ? 2 : 4);
  fn<<<(unsigned)(K * (chunks / cpc)), kThreads, K_::kSmemBytes, st>>>(maps, pv, rois, aligned, mp, lv, out, chunks, cpc);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

}  // namespace fwdr

// true when the row-streaming kernel can take the call (checked by cpm_roi_align_forward)
bool fwd_rows_supported(const cpm_pyramid_t* feat, int pooled_h, int pooled_w, int sampling_ratio, const void* d_out) {
  if (feat->layout != CPM_LAYOUT_NHWC || feat->dtype != CPM_F32) return false;
  if (pooled_h != pooled_w || (pooled_h != 7 && pooled_h != 14)) return false;
  if (sampling_ratio != 1 && sampling_ratio != 2) return false;
  const int ch = pooled_h == 7 ? 128 : 64;
  if (feat->channels % ch != 0 || feat->num_levels > fwdr::kMaxMapLevels) return false;
  if (((uintptr_t)d_out & 15) != 0) return false;
  for (int l = 0; l < feat->num_levels; l++)
    if (((uintptr_t)feat->d_level[l] & 15) != 0) return false;
  return fwdr::encode_fn() != nullptr;
}

int launch_fwd_rows(const cpm_pyramid_t* feat, const PyramidView& pv, const float* rois, long K, int P, int G, int aligned,
                    const MapperView& mp, const int* lv, float* out, cudaStream_t st) {
  const int CH = P == 7 ? 128 : 64;
  const fwdr::Maps* pm = nullptr;
  if (int rc = fwdr::pyramid_maps(&pm, feat, CH)) return rc;
  const fwdr::Maps& maps = *pm;
  if (P == 7) return G == 1 ? fwdr::launch_t<1, 1>(maps, pv, rois, K, aligned, mp, lv, out, st)
                            : fwdr::launch_t<1, 2>(maps, pv, rois, K, aligned, mp, lv, out, st);
  return G == 1 ? fwdr::launch_t<2, 1>(maps, pv, rois, K, aligned, mp, lv, out, st)
                : fwdr::launch_t<2, 2>(maps, pv, rois, K, aligned, mp, lv, out, st);
}

}  // namespace cpm
