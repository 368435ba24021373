// RoIAlign backward for the two CPM poolers (7x7 and 14x14, sampling_ratio 2) on sm_100a: deterministic, atomic-free,
// fed by the TMA, synchronised by mbarriers only.
//
// Replaces RoIAlignBackwardFeature (pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:259-365; host side :428-487), which
// scatters every sample's four taps with atomicAdd (:340-347) into an at::zeros map, once per FPN level.
//
// Decomposition ("pixel tiles own their gradient"):
//   bwd_prepare (roi_align_bwd.cu)  per-RoI sample taps (TapS), reach box, per 8-row band the bin rows with a tap in the
//                                   band (row clip); RoIs binned by (level, image) in RoI order
//   bwd_tile_lists                  one warp per 8x32-pixel tile: the STAGES of the tile, in RoI order -- for every RoI of
//                                   its (level, image) that reaches the tile, the clipped bin rows in groups of <= 4 (14x14)
//                                   / all of them (7x7).  A tile CTA starts from one list: no scan, no clipping.
//   bwd_tiles_tma                   one CTA (8 equal warps) per (tile, 64-channel chunk), a 5-slot ring of stages.
//     issue (warp k % 8 for stage k, three stages ahead): one cp.async.bulk of the RoI's taps and one
//       cp.async.bulk.tensor (14x14: a [64 ch][<= 4 bin rows] box of the (PH*PW, C, K) view of grad_out; UTMALDG) or
//       cp.async.bulk (7x7: the RoI's whole [64 ch][49] block; UBLKCP) -- grad_out stays (K, C, PH, PW) as the reference
//       lays it out, NO transposition -- completing on the slot's `full` mbarrier.
//     tables (all warps, one stage ahead, an eighth each): the separable weights of the stage from the taps in shared
//       memory -- WY[bin row][tile row]; per pixel column the weights of the bin columns -- then the `tab` mbarrier.
//     arithmetic: warp = (32-channel group, 8-column group), LANE = CHANNEL, thread = 8 rows x 8 columns of accumulators.
//       Because a lane owns a channel, the (channel, bin) order the copies deliver is read as is: the channel pitch in
//       shared memory is an odd number of 16-byte units (14x14, LDS.128) or of words (7x7, LDS.32) -- conflict-free.
//       Per bin row: horizontal pass h[x] = sum_q WX[q][x] * S[c][p][q] (packed FFMA2; weights are warp-uniform
//       broadcasts), then vertical pass acc[y][x] += WY[p][y] * h[x]; then the slot's `empty` mbarrier.
//     No block-wide barrier after the prologue.  The summation order is fixed (RoI index, bin row, bin column):
//     bit-identical run to run.
//   Every pixel of every level is written exactly once (zeros where nothing reaches): no memset pass.  The gradient is
//   written NHWC or NCHW (a lane owns a channel and 8 consecutive columns, so both are sector-aligned runs).
#include <cuda.h>

#include "roi_align_bwd.cuh"

namespace cpm {
namespace btma {

constexpr int NS = 5;              // stage ring
constexpr int LOOK = 3;            // stages in flight ahead of the arithmetic (slot reuse waits for stage k - NS + LOOK ... )
constexpr int NW = 8;              // warps: (channel group 0..1) x (column group 0..3)
constexpr int kThreads = NW * 32;
constexpr int G = 2;               // sampling grid per bin and axis

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

template <int PC>
struct XTab;
template <>
struct __align__(16) XTab<14> {
  // WT[a][x][t] = weight of bin column q = t - 2a on pixel column x0 + x (zero outside 0 <= q < 14): the two 16-byte
  // alignments a row of 14 bins can have inside a box that starts at a multiple of 4 bins
  float WT[2][TW][16];
  int krange[2][4];   // per (alignment, column group): first | last << 8 chunk of a row with weight on the group
};
template <>
struct __align__(16) XTab<7> {
  float W[7][TW];     // W[q][x]
  int qrange[4];      // per column group: first | last << 8 bin column with weight on the group (first > last: none)
};

template <int PC>
struct __align__(128) Smem {
  static constexpr int RMAX = PC == 14 ? 4 : 7;                                    // bin rows per stage
  static constexpr int STAGE_BYTES = PC == 14 ? 15 * 16 * CH : CH * 49 * 4;        // 15360 / 12544
  static constexpr int NT = 2 * PC * G;                                            // taps per RoI (y taps first)
  unsigned char S[NS][STAGE_BYTES];
  TapS taps[NS][NT];
  XTab<PC> tab[NS];
  float2 WY[NS][RMAX][TH];         // (w, w)
  int2 ent[NS];                    // the stage: {RoI, first bin row | rows << 8}
  uint64_t full[NS], tabb[NS], empty[NS];
};

// ---- per-tile stage lists ----------------------------------------------------------------------------------------------
// lists: segment-major; tile t of segment s owns SUB * n_s slots at SUB * sum_{tiles before t} n_{seg(tile)}  (no atomics:
// the layout is a function of seg_count alone).  Entry = {RoI, p0 | nr << 8}.
template <int PC>
__global__ void __launch_bounds__(256) bwd_tile_lists(PyramidView pv, TileGrid tg, int K, const int4* __restrict__ box,
                                                       const int* __restrict__ rowclip, int NB,
                                                       const int* __restrict__ seg_count, const int* __restrict__ perm,
                                                       int* __restrict__ tile_count, int* __restrict__ tile_off,
                                                       int2* __restrict__ lists, int ntiles) {
  constexpr int RMAX = Smem<PC>::RMAX;
  constexpr int SUB = (PC + RMAX - 1) / RMAX;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= ntiles) return;
  const TileId id = decode_tile(tg, pv.num_levels, t, TH, TW);
  long off = 0;
  for (int oi = 0; oi < pv.num_levels; oi++) {
    const int l = tg.order[oi];
    const int per = tg.tiles_x[l] * tg.tiles_y[l];
    if (l == id.l) {
      for (int b = 0; b < id.b; b++) off += (long)per * seg_count[l * pv.batch + b];
      off += (long)id.in_img * seg_count[l * pv.batch + id.b];
      break;
    }
    for (int b = 0; b < pv.batch; b++) off += (long)per * seg_count[l * pv.batch + b];
  }
  off *= SUB;
  const int seg = id.l * pv.batch + id.b;
  const int n = seg_count[seg];
  const int* plist = perm + (long)seg * K;
  const int band = id.y0 / TH;
  int2* out = lists + off;
  int cnt = 0;
  for (int base = 0; base < n; base += 128) {
    // four independent (list -> box -> row clip) chains per lane
    int r[4], rc[4];
    int4 bx[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = base + 32 * u + lane;
      r[u] = i < n ? plist[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) bx[u] = r[u] >= 0 ? __ldg(box + r[u]) : make_int4(1, 0, 1, 0);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const bool hit = bx[u].x <= bx[u].y && bx[u].x < id.y0 + TH && bx[u].y >= id.y0 && bx[u].z < id.x0 + TW && bx[u].w >= id.x0;
      rc[u] = hit && band < NB ? __ldg(rowclip + (long)r[u] * NB + band) : 1;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (base + 32 * u >= n) break;
      const int pA = rc[u] & 255, pB = rc[u] >> 8;
      const int ns = pA <= pB ? (pB - pA + RMAX) / RMAX : 0;
      int incl = ns;                                   // inclusive warp scan
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      int pos = cnt + incl - ns;
      for (int p0 = pA; p0 <= pB; p0 += RMAX) out[pos++] = make_int2(r[u], p0 | (min(RMAX, pB - p0 + 1) << 8));
      cnt += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (lane == 0) {
    tile_count[t] = cnt;
    tile_off[t] = (int)off;
  }
}

// ---- the tile kernel ------------------------------------------------------------------------------------------------
template <int PC, bool NCHW_OUT>
__global__ void __maxnreg__(112)
bwd_tiles_tma(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
              const __grid_constant__ CUtensorMap tm3, const __grid_constant__ CUtensorMap tm4, PyramidView pv, TileGrid tg,
              const float* __restrict__ go, const TapS* __restrict__ taps, const int* __restrict__ tile_count,
              const int* __restrict__ tile_off, const int2* __restrict__ lists, int chunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<PC>& sm = *reinterpret_cast<Smem<PC>*>(smem_raw);
  constexpr int NT = Smem<PC>::NT;
  const int C = pv.channels;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x / chunks;
  const int c0 = (blockIdx.x - t * chunks) * CH;
  const TileId id = decode_tile(tg, pv.num_levels, t, TH, TW);
  const int y0 = id.y0, x0 = id.x0;
  const int n = tile_count[t];
  const int2* ent = lists + tile_off[t];

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.tabb[s], NW);
      mbar_init(&sm.empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (PC == 14) {
    // the constant zero padding of the x tables (q = 14, 15 at alignment 0; q = -2, -1 at alignment 1)
    for (int e = threadIdx.x; e < NS * TW; e += kThreads) {
      XTab<14>& tb = sm.tab[e / TW];
      const int x = e % TW;
      tb.WT[0][x][14] = tb.WT[0][x][15] = 0.f;
      tb.WT[1][x][0] = tb.WT[1][x][1] = 0.f;
    }
  }
  __syncthreads();

  const int cg = warp & 1, g = warp >> 1;          // channel group, column group
  const int chl = 32 * cg + lane;                  // channel inside the CTA's chunk

  // ---- the three roles of a warp ----
  int2 mye = make_int2(0, 0);                      // entry of the next stage this warp issues (k = warp, warp + 8, ...)
  if (warp < n) mye = __ldg(ent + warp);
  auto issue = [&](int k) {                        // k % NW == warp
    const int s = k % NS;
    if (k >= NS) mbar_wait(&sm.empty[s], ((k / NS) - 1) & 1);
    const int2 e = mye;
    if (k + NW < n) mye = __ldg(ent + k + NW);
    if (lane == 0) {
      sm.ent[s] = e;
      const int r = e.x, p0 = e.y & 255, nr = e.y >> 8;
      if constexpr (PC == 14) {
        const CUtensorMap* tm = nr == 1 ? &tm1 : nr == 2 ? &tm2 : nr == 3 ? &tm3 : &tm4;
        const int nck = nr == 1 ? 5 : nr == 2 ? 9 : nr == 3 ? 11 : 15;
        mbar_expect_tx(&sm.full[s], (uint32_t)(nck * 16 * CH + NT * 16));
        bulk_g2s(&sm.taps[s][0], taps + (long)r * NT, NT * 16, &sm.full[s]);
        tma_load_3d(&sm.S[s][0], tm, (14 * p0) & ~3, c0, r, &sm.full[s]);
      } else {
        mbar_expect_tx(&sm.full[s], (uint32_t)(CH * 49 * 4 + NT * 16));
        bulk_g2s(&sm.taps[s][0], taps + (long)r * NT, NT * 16, &sm.full[s]);
        bulk_g2s(&sm.S[s][0], go + ((long)r * C + c0) * 49, (uint32_t)(CH * 49 * 4), &sm.full[s]);
      }
    }
  };
  auto build = [&](int k) {                        // this warp's eighth of the stage's weight tables
    const int s = k % NS;
    mbar_wait(&sm.full[s], (k / NS) & 1);
    const TapS* tY = &sm.taps[s][0];
    const TapS* tX = tY + PC * G;
    XTab<PC>& tb = sm.tab[s];
    const int e = sm.ent[s].y;
    const int p0 = e & 255, nr = e >> 8;
    if constexpr (PC == 14) {
      // x weights: pixel columns 4 * warp .. + 3, lane = (column, bin-column pair)
      const int x = 4 * warp + (lane >> 3), j = lane & 7;
      if (j < 7) {
        const int px = x0 + x;
        const float w0 = 0.5f * (tap_weight(tX[4 * j], px) + tap_weight(tX[4 * j + 1], px));
        const float w1 = 0.5f * (tap_weight(tX[4 * j + 2], px) + tap_weight(tX[4 * j + 3], px));
        *reinterpret_cast<float2*>(&tb.WT[0][x][2 * j]) = make_float2(w0, w1);
        *reinterpret_cast<float2*>(&tb.WT[1][x][2 * j + 2]) = make_float2(w0, w1);
      }
    } else {
      const int x = 4 * warp + lane / 7, q = lane % 7;
      if (lane < 28) {
        const int px = x0 + x;
        tb.W[q][x] = 0.5f * (tap_weight(tX[2 * q], px) + tap_weight(tX[2 * q + 1], px));
      }
    }
    if (warp < 4) {
      // bin columns with a tap on column group `warp`
      bool hg = false;
      if (lane < PC * G) {
        const TapS tq = tX[lane];
        hg = tq.lo >= 0 && tq.hi >= x0 + 8 * warp && tq.lo < x0 + 8 * warp + 8;
      }
      const unsigned mg = __ballot_sync(0xffffffffu, hg);
      if (lane == 0) {
        if constexpr (PC == 14) {
          if (mg) {
            const int qa = (__ffs(mg) - 1) / G, qb = (31 - __clz(mg)) / G;
            tb.krange[0][warp] = (qa >> 2) | ((qb >> 2) << 8);
            tb.krange[1][warp] = ((qa + 2) >> 2) | (((qb + 2) >> 2) << 8);
          } else {
            tb.krange[0][warp] = tb.krange[1][warp] = 1;     // first 1 > last 0
          }
        } else {
          tb.qrange[warp] = mg ? ((__ffs(mg) - 1) / G) | (((31 - __clz(mg)) / G) << 8) : 1;
        }
      }
    }
    if (warp < nr && lane < TH) {
      // y weights of bin row p0 + warp on the 8 tile rows
      const int p = p0 + warp;
      const float w = 0.5f * (tap_weight(tY[2 * p], y0 + lane) + tap_weight(tY[2 * p + 1], y0 + lane));
      sm.WY[s][warp][lane] = make_float2(w, w);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.tabb[s]);
  };

  u64 acc[TH][4];
#pragma unroll
  for (int y = 0; y < TH; y++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[y][j] = 0ull;

  auto vertical = [&](const u64 (&hp)[4], const u64* wy) {
#pragma unroll
    for (int y = 0; y < TH; y++) {
      const u64 w = wy[y];
#pragma unroll
      for (int j = 0; j < 4; j++) acc[y][j] = fma2(w, hp[j], acc[y][j]);
    }
  };
  auto compute = [&](int k) {
    const int s = k % NS;
    mbar_wait(&sm.tabb[s], (k / NS) & 1);
    const int e = sm.ent[s].y;
    const int p0 = e & 255, nr = e >> 8;
    const XTab<PC>& tb = sm.tab[s];
    if constexpr (PC == 14) {
      if ((tb.krange[0][g] & 255) <= (tb.krange[0][g] >> 8)) {
        const int b0 = (14 * p0) & ~3;
        const int pitch = 16 * (nr == 1 ? 5 : nr == 2 ? 9 : nr == 3 ? 11 : 15);
        const unsigned char* Sc = &sm.S[s][0] + chl * pitch;
        for (int pi = 0; pi < nr; pi++) {
          const int eo = 14 * (p0 + pi) - b0;
          const int a = (eo >> 1) & 1;
          const int kr = tb.krange[a][g];
          const int k0 = kr & 255, k1 = kr >> 8;
          u64 h2[8];
#pragma unroll
          for (int x = 0; x < 8; x++) h2[x] = 0ull;
          const ulonglong2* Sp = reinterpret_cast<const ulonglong2*>(Sc + 16 * ((eo >> 2) + k0));
          const ulonglong2* Wp = reinterpret_cast<const ulonglong2*>(&tb.WT[a][8 * g][4 * k0]);
          for (int kk = k0; kk <= k1; kk++, Sp++, Wp++) {
            const ulonglong2 sv = *Sp;
#pragma unroll
            for (int x = 0; x < 8; x++) {
              const ulonglong2 w = Wp[4 * x];
              h2[x] = fma2(sv.x, w.x, h2[x]);
              h2[x] = fma2(sv.y, w.y, h2[x]);
            }
          }
          u64 hp[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float a0, a1, b0f, b1f;
            unpack2(h2[2 * j], a0, a1);
            unpack2(h2[2 * j + 1], b0f, b1f);
            hp[j] = pack2(__fadd_rn(a0, a1), __fadd_rn(b0f, b1f));
          }
          vertical(hp, reinterpret_cast<const u64*>(&sm.WY[s][pi][0]));
        }
      }
    } else {
      const int qr = tb.qrange[g];
      const int q0 = qr & 255, q1 = qr >> 8;
      if (q0 <= q1) {
        const float* Sc = reinterpret_cast<const float*>(&sm.S[s][0]) + chl * 49;
        for (int pi = 0; pi < nr; pi++) {
          u64 hp[4] = {0ull, 0ull, 0ull, 0ull};
          const float* Sp = Sc + (p0 + pi) * 7 + q0;
          const ulonglong2* Wp = reinterpret_cast<const ulonglong2*>(&tb.W[q0][8 * g]);
          for (int q = q0; q <= q1; q++, Sp++, Wp += TW / 4) {
            const float v = *Sp;
            const u64 v2 = pack2(v, v);
            const ulonglong2 w0 = Wp[0], w1 = Wp[1];
            hp[0] = fma2(v2, w0.x, hp[0]);
            hp[1] = fma2(v2, w0.y, hp[1]);
            hp[2] = fma2(v2, w1.x, hp[2]);
            hp[3] = fma2(v2, w1.y, hp[3]);
          }
          vertical(hp, reinterpret_cast<const u64*>(&sm.WY[s][pi][0]));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  };

  // ---- the pipeline: stage i + LOOK is issued, the tables of stage i + 1 are built, stage i is accumulated ----
  for (int k = 0; k < LOOK && k < n; k++)
    if (k % NW == warp) issue(k);
  if (n > 0) build(0);
  for (int i = 0; i < n; i++) {
    const int k = i + LOOK;
    if (k < n && k % NW == warp) issue(k);
    if (i + 1 < n) build(i + 1);
    compute(i);
  }

  // ---- the tile's gradient: every pixel written exactly once ----
  const int H = pv.H[id.l], W = pv.W[id.l];
  const int c = c0 + chl;
  const int xg = x0 + 8 * g;
  if (NCHW_OUT) {
    float* dst = (float*)pv.ptr[id.l] + (((long)id.b * C + c) * H + y0) * W + xg;
    const bool v4 = (W & 3) == 0 && xg + 8 <= W;
    const bool v8 = (W & 7) == 0 && xg + 8 <= W;
#pragma unroll
    for (int y = 0; y < TH; y++) {
      if (y0 + y >= H) break;
      float* row = dst + (long)y * W;
      if (v8) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 4; j++) unpack2(acc[y][j], a[2 * j], a[2 * j + 1]);
        st_global_v8(row, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
      } else if (v4) {
        reinterpret_cast<ulonglong2*>(row)[0] = make_ulonglong2(acc[y][0], acc[y][1]);
        reinterpret_cast<ulonglong2*>(row)[1] = make_ulonglong2(acc[y][2], acc[y][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float lo, hi;
          unpack2(acc[y][j], lo, hi);
          if (xg + 2 * j < W) row[2 * j] = lo;
          if (xg + 2 * j + 1 < W) row[2 * j + 1] = hi;
        }
      }
    }
  } else {
    float* dst = (float*)pv.ptr[id.l] + (((long)id.b * H + y0) * W + xg) * C + c;
#pragma unroll
    for (int y = 0; y < TH; y++) {
      if (y0 + y >= H) break;
      float* row = dst + (long)y * W * C;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float lo, hi;
        unpack2(acc[y][j], lo, hi);
        if (xg + 2 * j < W) row[(long)(2 * j) * C] = lo;
        if (xg + 2 * j + 1 < W) row[(long)(2 * j + 1) * C] = hi;
      }
    }
  }
}

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

long num_tiles(const cpm_pyramid_t* p) {
  TileGrid tg;
  return make_tile_grid(tg, p, TH, TW);
}

size_t list_entries(const cpm_pyramid_t* p, int64_t K, int P) {
  long mx = 1;
  for (int l = 0; l < p->num_levels; l++) {
    const long per = (long)((p->width[l] + TW - 1) / TW) * ((p->height[l] + TH - 1) / TH);
    if (per > mx) mx = per;
  }
  const int sub = P == 14 ? 4 : 1;
  return (size_t)mx * (size_t)(K > 0 ? K : 1) * sub;
}

template <int PC, bool NCHW_OUT>
static int launch_tiles(const CUtensorMap* tm, const PyramidView& pv, const TileGrid& tg, const float* go, const TapS* taps,
                        const int* tile_count, const int* tile_off, const int2* lists, long tiles, int chunks,
                        cudaStream_t st) {
  auto fn = bwd_tiles_tma<PC, NCHW_OUT>;
  static thread_local int configured_dev = -1;
  int dev;
  CPM_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CPM_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<PC>)));
    configured_dev = dev;
  }
  fn<<<(unsigned)(tiles * chunks), kThreads, sizeof(Smem<PC>), st>>>(tm[0], tm[1], tm[2], tm[3], pv, tg, go, taps, tile_count,
                                                                      tile_off, lists, chunks);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

int launch(const cpm_pyramid_t* grad_feat, const PyramidView& pv, const float* go, int K, int P, const TapS* taps,
           const int4* box, const int* rowclip, int NB, const int* seg_count, const int* perm, int* tile_count,
           int* tile_off, int2* lists, cudaStream_t st) {
  const int C = grad_feat->channels;
  TileGrid tg;
  const long tiles = make_tile_grid(tg, grad_feat, TH, TW);
  const int chunks = C / CH;
  CPM_CHECK_ARG(tiles * chunks < (1L << 31), "gradient pyramid too large for one launch");
  CUtensorMap tm[4];
  memset(tm, 0, sizeof(tm));
  if (P == 14 && K > 0) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CPM_ERR_UNSUPPORTED;
    }
    const int nck[4] = {5, 9, 11, 15};
    for (int i = 0; i < 4; i++) {
      const cuuint64_t gdim[3] = {196, (cuuint64_t)C, (cuuint64_t)K};
      const cuuint64_t gstr[2] = {196 * 4, (cuuint64_t)C * 196 * 4};
      const cuuint32_t box3[3] = {(cuuint32_t)(4 * nck[i]), (cuuint32_t)CH, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUresult rc = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)go, gdim, gstr, box3, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc);
        return CPM_ERR_UNSUPPORTED;
      }
    }
  }
  if (P == 14)
    bwd_tile_lists<14><<<(unsigned)((tiles + 7) / 8), 256, 0, st>>>(pv, tg, K > 0 ? K : 1, box, rowclip, NB, seg_count, perm,
                                                                   tile_count, tile_off, lists, (int)tiles);
  else
    bwd_tile_lists<7><<<(unsigned)((tiles + 7) / 8), 256, 0, st>>>(pv, tg, K > 0 ? K : 1, box, rowclip, NB, seg_count, perm,
                                                                  tile_count, tile_off, lists, (int)tiles);
  CPM_CHECK_LAUNCH();
  const bool nchw = grad_feat->layout == CPM_LAYOUT_NCHW;
  if (P == 14)
    return nchw ? launch_tiles<14, true>(tm, pv, tg, go, taps, tile_count, tile_off, lists, tiles, chunks, st)
                : launch_tiles<14, false>(tm, pv, tg, go, taps, tile_count, tile_off, lists, tiles, chunks, st);
  return nchw ? launch_tiles<7, true>(tm, pv, tg, go, taps, tile_count, tile_off, lists, tiles, chunks, st)
              : launch_tiles<7, false>(tm, pv, tg, go, taps, tile_count, tile_off, lists, tiles, chunks, st);
}

}  // namespace btma
}  // namespace cpm
