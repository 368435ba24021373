// Definitions shared by the deterministic RoIAlign backward kernels (roi_align_bwd.cu: prepare pass + the generic staged
// tile kernel; roi_align_bwd_tma.cu: the warp-specialised TMA kernel of the two CPM poolers).
#pragma once
#include "common.cuh"

namespace cpm {

typedef unsigned long long u64;

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
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// one full 32-byte sector per lane and instruction (STG.256, sm_100): what an NCHW write-out of 8 consecutive columns wants
__device__ __forceinline__ void st_global_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f),
               "f"(g), "f"(h)
               : "memory");
}

// One sample coordinate of one RoI along one axis (bilinear_interpolate_gradient, ROIAlign_cuda.cu:113-171, per axis).
struct __align__(16) TapS {
  int lo, hi;       // lo < 0: sample out of range
  float wlo, whi;   // (1 - l), l
};

__device__ __forceinline__ TapS ld_tap(const TapS* p) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  TapS t;
  t.lo = v.x;
  t.hi = v.y;
  t.wlo = __int_as_float(v.z);
  t.whi = __int_as_float(v.w);
  return t;
}

__device__ __forceinline__ float tap_weight(const TapS& t, int pix) {
  return (t.lo == pix ? t.wlo : 0.f) + (t.hi == pix ? t.whi : 0.f);
}

// Pixel tiles of the gradient pyramid in launch order (coarsest level first: its tiles carry the longest RoI lists).
struct TileGrid {
  int tiles_x[CPM_MAX_LEVELS], tiles_y[CPM_MAX_LEVELS];
  int first[CPM_MAX_LEVELS + 1];       // first tile id of the level in launch order
  int order[CPM_MAX_LEVELS];           // launch order -> level
};

struct TileId {
  int l, b, y0, x0, in_img, per_img;
};

__device__ __forceinline__ TileId decode_tile(const TileGrid& tg, int num_levels, int t, int th, int tw) {
  int oi = 0;
  while (oi + 1 < num_levels && t >= tg.first[oi + 1]) oi++;
  TileId id;
  id.l = tg.order[oi];
  t -= tg.first[oi];
  id.per_img = tg.tiles_x[id.l] * tg.tiles_y[id.l];
  id.b = t / id.per_img;
  id.in_img = t - id.b * id.per_img;
  id.y0 = (id.in_img / tg.tiles_x[id.l]) * th;
  id.x0 = (id.in_img % tg.tiles_x[id.l]) * tw;
  return id;
}

inline long make_tile_grid(TileGrid& tg, const cpm_pyramid_t* p, int th, int tw) {
  const int L = p->num_levels;
  long tiles = 0;
  for (int i = 0; i < L; i++) {
    const int l = L - 1 - i;
    tg.order[i] = l;
    tg.tiles_x[l] = (p->width[l] + tw - 1) / tw;
    tg.tiles_y[l] = (p->height[l] + th - 1) / th;
    tg.first[i] = (int)tiles;
    tiles += (long)p->batch * tg.tiles_x[l] * tg.tiles_y[l];
  }
  for (int i = L; i <= CPM_MAX_LEVELS; i++) tg.first[i] = (int)tiles;
  for (int i = L; i < CPM_MAX_LEVELS; i++) tg.order[i] = 0, tg.tiles_x[i] = tg.tiles_y[i] = 0;
  return tiles;
}

// ---- the TMA kernel (roi_align_bwd_tma.cu) ----
namespace btma {
constexpr int TH = 8, TW = 32;     // pixel tile of one CTA
constexpr int CH = 64;             // channels of one CTA
// the pooler shapes the kernel is built for
inline bool shape_ok(int PH, int PW, int G) { return G == 2 && PH == PW && (PH == 7 || PH == 14); }
// capacity of the per-tile stage lists: sum over (level, image) of tiles * RoIs of that segment * stages per (tile, RoI)
// <= max tiles of a level * K * ceil(P / rows per stage)
size_t list_entries(const cpm_pyramid_t* p, int64_t K, int P);
long num_tiles(const cpm_pyramid_t* p);
// builds the per-tile stage lists (after bwd_prepare) and runs the tile kernel; returns CPM_ERR_UNSUPPORTED when the
// driver cannot encode the tensor maps (the caller then takes the generic staged kernel)
int launch(const cpm_pyramid_t* grad_feat, const PyramidView& pv, const float* go, int K, int P, const TapS* taps,
           const int4* box, const int* rowclip, int NB, const int* seg_count, const int* perm, int* tile_count,
           int* tile_off, int2* lists, cudaStream_t st);
}  // namespace btma

}  // namespace cpm
