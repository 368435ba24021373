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
