// NCHW <-> NHWC staging of one feature-map level.
//
// The reference's kernels index (B, C, H, W) maps directly (ROIAlign_cuda.cu:219); the B200 hot kernels want the
// channel vector of a pixel contiguous.  A torch.channels_last tensor already is; anything else is staged once with
// this tiled transpose (per image: a C x HW matrix -> HW x C), 32x32 tiles through padded shared memory, every global
// access a full 128-byte row.  Pure HBM streaming: 2 * B*C*H*W*elsize bytes.
#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

// in: (B, R, S) -> out: (B, S, R)
template <typename T>
__global__ void __launch_bounds__(256) transpose_tiles(const T* __restrict__ in, T* __restrict__ out, int R, int S) {
  __shared__ T tile[32][33];
  const long img = blockIdx.z;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* src = in + img * (long)R * S;
  T* dst = out + img * (long)R * S;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int r = r0 + ty + 8 * k, s = s0 + tx;
    if (r < R && s < S) tile[ty + 8 * k][tx] = src[(long)r * S + s];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int s = s0 + ty + 8 * k, r = r0 + tx;
    if (r < R && s < S) dst[(long)s * R + r] = tile[tx][ty + 8 * k];
  }
}


// fp32 pyramid, all levels in one launch, 16-byte accesses on both sides.  A CTA moves one 64-channel x 64-pixel tile of
// one image of one level: it reads 4 consecutive pixels of a channel per thread (256-byte runs of a channel plane), and
// writes 4 consecutive channels of a pixel per thread (256-byte runs of a pixel's channel vector), through a padded
// shared-memory tile.  to_nhwc = false runs the same tile the other way.
struct StageLevels {
  const float* src[CPM_MAX_LEVELS];
  float* dst[CPM_MAX_LEVELS];
  int hw[CPM_MAX_LEVELS];
  int first[CPM_MAX_LEVELS + 1];   // first CTA of the level
  int ptiles[CPM_MAX_LEVELS];      // pixel tiles per image
  int num_levels, batch, channels, ctiles;
};

__global__ void __launch_bounds__(256) stage_pyramid_f32(StageLevels sl, int to_nhwc) {
  __shared__ float tile[64][65];
  int l = 0;
  while (l + 1 < sl.num_levels && (int)blockIdx.x >= sl.first[l + 1]) l++;
  int t = blockIdx.x - sl.first[l];
  const int C = sl.channels, HW = sl.hw[l];
  const int ct = t % sl.ctiles;
  t /= sl.ctiles;
  const int pt = t % sl.ptiles[l];
  const int img = t / sl.ptiles[l];
  const int c0 = ct * 64, p0 = pt * 64;
  const float* cmaj = (to_nhwc ? sl.src[l] : sl.dst[l]) + (long)img * C * HW;    // (C, HW) side
  const float* pmaj = (to_nhwc ? sl.dst[l] : sl.src[l]) + (long)img * C * HW;    // (HW, C) side
  const int a = threadIdx.x >> 4, b4 = (threadIdx.x & 15) * 4;
  const bool pvec = (HW & 3) == 0, cvec = (C & 3) == 0;
  if (to_nhwc) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = c0 + a + 16 * k, p = p0 + b4;
      if (c < C) {
        const float* s = cmaj + (long)c * HW + p;
        if (pvec && p + 3 < HW) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(s));
          tile[a + 16 * k][b4] = v.x; tile[a + 16 * k][b4 + 1] = v.y; tile[a + 16 * k][b4 + 2] = v.z; tile[a + 16 * k][b4 + 3] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (p + i < HW) tile[a + 16 * k][b4 + i] = __ldg(s + i);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p = p0 + a + 16 * k, c = c0 + b4;
      if (p < HW && c < C) {
        float* d = const_cast<float*>(pmaj) + (long)p * C + c;
        const float v0 = tile[b4][a + 16 * k], v1 = tile[b4 + 1][a + 16 * k], v2 = tile[b4 + 2][a + 16 * k], v3 = tile[b4 + 3][a + 16 * k];
        if (cvec && c + 3 < C) {
          *reinterpret_cast<float4*>(d) = make_float4(v0, v1, v2, v3);
        } else {
          d[0] = v0;
          if (c + 1 < C) d[1] = v1;
          if (c + 2 < C) d[2] = v2;
          if (c + 3 < C) d[3] = v3;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p = p0 + a + 16 * k, c = c0 + b4;
      if (p < HW && c < C) {
        const float* s = pmaj + (long)p * C + c;
        if (cvec && c + 3 < C) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(s));
          tile[b4][a + 16 * k] = v.x; tile[b4 + 1][a + 16 * k] = v.y; tile[b4 + 2][a + 16 * k] = v.z; tile[b4 + 3][a + 16 * k] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (c + i < C) tile[b4 + i][a + 16 * k] = __ldg(s + i);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = c0 + a + 16 * k, p = p0 + b4;
      if (c < C) {
        float* d = const_cast<float*>(cmaj) + (long)c * HW + p;
        const float v0 = tile[a + 16 * k][b4], v1 = tile[a + 16 * k][b4 + 1], v2 = tile[a + 16 * k][b4 + 2], v3 = tile[a + 16 * k][b4 + 3];
        if (pvec && p + 3 < HW) {
          *reinterpret_cast<float4*>(d) = make_float4(v0, v1, v2, v3);
        } else {
          if (p < HW) d[0] = v0;
          if (p + 1 < HW) d[1] = v1;
          if (p + 2 < HW) d[2] = v2;
          if (p + 3 < HW) d[3] = v3;
        }
      }
    }
  }
}

int check_pyramid(const cpm_pyramid_t* p, const char* what);

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_layout_convert(const void* d_src, void* d_dst, int batch, int channels, int height, int width,
                                  int dtype, int to_layout, void* stream) {
  CPM_CHECK_ARG(batch >= 0 && channels >= 1 && height >= 1 && width >= 1, "bad shape");
  CPM_CHECK_ARG(to_layout == CPM_LAYOUT_NCHW || to_layout == CPM_LAYOUT_NHWC, "unknown layout %d", to_layout);
  CPM_CHECK_ARG(batch < 65536, "batch too large");
  if (batch == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_src, "src")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_dst, "dst")) != CPM_OK) return rc;
  const int HW = height * width;
  // NCHW -> NHWC: rows = C, cols = HW ; NHWC -> NCHW: rows = HW, cols = C
  const int R = to_layout == CPM_LAYOUT_NHWC ? channels : HW;
  const int S = to_layout == CPM_LAYOUT_NHWC ? HW : channels;
  dim3 grid((S + 31) / 32, (R + 31) / 32, batch);
  CPM_CHECK_ARG(grid.y < 65536, "map too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CPM_F32)
    transpose_tiles<float><<<grid, 256, 0, st>>>((const float*)d_src, (float*)d_dst, R, S);
  else if (dtype == CPM_F64)
    transpose_tiles<double><<<grid, 256, 0, st>>>((const double*)d_src, (double*)d_dst, R, S);
  else if (dtype == CPM_BF16)
    transpose_tiles<unsigned short><<<grid, 256, 0, st>>>((const unsigned short*)d_src, (unsigned short*)d_dst, R, S);
  else {
    set_error("unknown dtype %d", dtype);
    return CPM_ERR_INVALID_ARG;
  }
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

extern "C" int cpm_layout_convert_pyramid(const cpm_pyramid_t* src, const cpm_pyramid_t* dst, void* stream) {
  int rc;
  if ((rc = check_pyramid(src, "src")) != CPM_OK) return rc;
  if ((rc = check_pyramid(dst, "dst")) != CPM_OK) return rc;
  CPM_CHECK_ARG(src->num_levels == dst->num_levels && src->batch == dst->batch && src->channels == dst->channels &&
                    src->dtype == dst->dtype,
                "src and dst pyramids differ in shape or dtype");
  CPM_CHECK_ARG(src->layout != dst->layout, "src and dst have the same layout");
  for (int l = 0; l < src->num_levels; l++)
    CPM_CHECK_ARG(src->height[l] == dst->height[l] && src->width[l] == dst->width[l], "level %d differs in size", l);
  if (src->batch == 0) return CPM_OK;
  bool fast = src->dtype == CPM_F32;
  for (int l = 0; fast && l < src->num_levels; l++)
    fast = (((uintptr_t)src->d_level[l] | (uintptr_t)dst->d_level[l]) & 15) == 0;
  if (!fast) {
    for (int l = 0; l < src->num_levels; l++)
      if ((rc = cpm_layout_convert(src->d_level[l], dst->d_level[l], src->batch, src->channels, src->height[l], src->width[l],
                                   src->dtype, dst->layout, stream)) != CPM_OK)
        return rc;
    return CPM_OK;
  }
  StageLevels sl;
  sl.num_levels = src->num_levels;
  sl.batch = src->batch;
  sl.channels = src->channels;
  sl.ctiles = (src->channels + 63) / 64;
  long ctas = 0;
  for (int l = 0; l < CPM_MAX_LEVELS; l++) {
    sl.first[l] = (int)ctas;
    if (l < src->num_levels) {
      sl.src[l] = (const float*)src->d_level[l];
      sl.dst[l] = (float*)dst->d_level[l];
      sl.hw[l] = src->height[l] * src->width[l];
      sl.ptiles[l] = (sl.hw[l] + 63) / 64;
      ctas += (long)src->batch * sl.ptiles[l] * sl.ctiles;
    } else {
      sl.src[l] = nullptr;
      sl.dst[l] = nullptr;
      sl.hw[l] = sl.ptiles[l] = 0;
    }
  }
  sl.first[CPM_MAX_LEVELS] = (int)ctas;
  CPM_CHECK_ARG(ctas < (1L << 31), "pyramid too large for one launch");
  stage_pyramid_f32<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(sl, dst->layout == CPM_LAYOUT_NHWC ? 1 : 0);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
