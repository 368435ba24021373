// Grid-point heat-map -> box decode on the device.
//
// Replaces GridPostProcessor.get_boxes (pet/rcnn/modeling/grid_cascade_rcnn/inference.py:189-279), which copies the
// whole (R, P, h, w) heat-map stack and the boxes to the host (:195-196), reduces there and copies the boxes back
// (:278).  Here one CTA decodes one RoI: a warp per grid point streams that point's h*w logits with 16-byte loads,
// keeps the running (sigmoid, first index) maximum per lane and finishes with a shuffle reduction; four threads then
// form the score-weighted boundary votes (:251-271).  Traffic = the logits once (R*P*h*w*4 bytes) + 32 bytes per RoI.
#include <stdlib.h>

#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

constexpr int kMaxPoints = 64;
constexpr int kRegVec = 7;             // float4 per lane of the register-resident fast path: maps up to 896 logits (28x28 = 784)

struct SubXY {
  int v[2 * kMaxPoints];
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void argmax_merge(float& bs, int& bi, float s, int i) {
  // torch.max(dim): the first index among equal maxima (inference.py:204)
  if (s > bs || (s == bs && i < bi)) {
    bs = s;
    bi = i;
  }
}

// Register-resident arg-max of one heat-map of n4 <= 32 * kRegVec float4 (see the comment at the call site): pass 1 finds the
// largest logit, pass 2 evaluates the sigmoid only where fp32 sigmoids can tie.  SMEM: the map is in shared memory.
__device__ __forceinline__ void warp_argmax(float& bs, int& bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_merge(bs, bi, s2, i2);
  }
}

// On return (bs, bi) is the same on every lane of the warp.
template <bool SMEM>
__device__ __forceinline__ void point_argmax_regs(const float4* __restrict__ m4, int n4, int lane, float& bs, int& bi) {
  float4 v[kRegVec];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kRegVec; k++) {
    const int i = lane + 32 * k;
    v[k] = i < n4 ? (SMEM ? m4[i] : __ldg(m4 + i)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
  }
  const float lmx = mx;                  // this lane's largest logit
  {
    // warp maximum with one redux.sync: floats ordered as signed integers after flipping the magnitude bits of negatives
    // (lmx is never NaN: fmaxf dropped them)
    int key = __float_as_int(lmx);
    key ^= (key >> 31) & 0x7fffffff;
    key = __reduce_max_sync(0xffffffffu, key);
    key ^= (key >> 31) & 0x7fffffff;
    mx = __int_as_float(key);
  }
  const float t = fminf(mx - 1e-3f, 8.0f);
  // (NaN logits fail `>= t` and fmaxf ignores them, so they never win -- as in the reference's `>` comparison; an
  //  all-NaN map leaves bi unset and decodes as index 0)
  // Census of the band [t, inf), branch-free and only on lanes that can have a member: how many logits, and the lane's
  // first one.  Nearly always the band holds exactly one logit -- the maximum itself -- and the answer is
  // (sigmoid(max), its index) without any per-element sigmoid or merge.
  int cnt = 0, first = 0x7fffffff;
  if (lmx >= t) {
#pragma unroll
    for (int k = kRegVec - 1; k >= 0; k--) {       // descending, so that `first` ends as the smallest index
      const int i = 4 * (lane + 32 * k);
      const bool hw_ = v[k].w >= t, hz = v[k].z >= t, hy = v[k].y >= t, hx = v[k].x >= t;
      cnt += (int)hw_ + (int)hz + (int)hy + (int)hx;
      first = hw_ ? i + 3 : first;
      first = hz ? i + 2 : first;
      first = hy ? i + 1 : first;
      first = hx ? i + 0 : first;
    }
  }
  const unsigned owners = __ballot_sync(0xffffffffu, cnt > 0);
  if (__popc(owners) == 1 && __shfl_sync(0xffffffffu, cnt, __ffs(owners) - 1) == 1) {
    bs = sigmoidf_ref(mx);               // (every lane holds the same pair: the caller's merge is then a no-op)
    bi = __shfl_sync(0xffffffffu, first, __ffs(owners) - 1);
    return;
  }
#pragma unroll
  for (int k = 0; k < kRegVec; k++) {
    const int i = 4 * (lane + 32 * k);
    if (v[k].x >= t) argmax_merge(bs, bi, sigmoidf_ref(v[k].x), i + 0);
    if (v[k].y >= t) argmax_merge(bs, bi, sigmoidf_ref(v[k].y), i + 1);
    if (v[k].z >= t) argmax_merge(bs, bi, sigmoidf_ref(v[k].z), i + 2);
    if (v[k].w >= t) argmax_merge(bs, bi, sigmoidf_ref(v[k].w), i + 3);
  }
  warp_argmax(bs, bi);
}

__global__ void __launch_bounds__(288, 4) grid_decode_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                           int P, int gs, int h, int w, SubXY sub, float ratio,
                                                           float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                           unsigned wmagic) {
  __shared__ float sc[kMaxPoints], ax[kMaxPoints], ay[kMaxPoints];
  const long r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = h * w;
  const float4 bx = *reinterpret_cast<const float4*>(boxes + 4 * r);
  const float width = bx.z - bx.x, height = bx.w - bx.y;
  const float x1 = bx.x - ratio * (width / 2), y1 = bx.y - ratio * (height / 2);
  for (int p = warp; p < P; p += nwarps) {
    const float* m = logits + (r * P + p) * (long)hw;
    float bs = -1.f;
    int bi = 0x7fffffff;
    if ((hw & 3) == 0 && hw <= 4 * 32 * kRegVec) {
      // Fast path: the map sits in registers (<= kRegVec float4 per lane, one streaming read).  sigmoid is monotone, so
      // the arg-max of the sigmoids is found on the logits: pass 1 takes the largest logit M; pass 2 evaluates the sigmoid
      // only where it could still equal (or, through rounding, exceed) sigmoid(M) -- logits within 1e-3 of M, and every
      // logit >= 8, where fp32 sigmoids start to collide (1 - e^-x within an ulp of its neighbours) up to saturating at
      // exactly 1.0 -- and applies the reference's rule there (largest sigmoid, first index among equals).  Below that
      // band a logit's sigmoid is smaller than sigmoid(M) by more than 4 ulp, so the result equals the full evaluation.
      point_argmax_regs<false>(reinterpret_cast<const float4*>(m), hw >> 2, lane, bs, bi);
    } else if ((hw & 3) == 0) {
      const float4* m4 = reinterpret_cast<const float4*>(m);
      for (int i = lane; i < hw / 4; i += 32) {
        const float4 v = __ldg(m4 + i);
        argmax_merge(bs, bi, sigmoidf_ref(v.x), 4 * i + 0);
        argmax_merge(bs, bi, sigmoidf_ref(v.y), 4 * i + 1);
        argmax_merge(bs, bi, sigmoidf_ref(v.z), 4 * i + 2);
        argmax_merge(bs, bi, sigmoidf_ref(v.w), 4 * i + 3);
      }
    } else {
      for (int i = lane; i < hw; i += 32) argmax_merge(bs, bi, sigmoidf_ref(__ldg(m + i)), i);
    }
    if (!((hw & 3) == 0 && hw <= 4 * 32 * kRegVec)) warp_argmax(bs, bi);   // (the register path returns warp-uniform values)
    if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;   // all-NaN map
      const int row = wmagic ? (int)(((unsigned)bi * wmagic) >> 16) : bi / w;
      const int xs = bi - row * w + sub.v[2 * p], ys = row + sub.v[2 * p + 1];
      sc[p] = bs;
      ax[p] = ((float)xs + 0.5f) / (float)(2 * w) * (1.f + ratio) * width + x1;      // inference.py:246
      ay[p] = ((float)ys + 0.5f) / (float)(2 * h) * (1.f + ratio) * height + y1;     // inference.py:247
      if (out_scores) out_scores[r * P + p] = bs;
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    // boundary votes, inference.py:251-271: x1 <- points [0..gs), y1 <- [0, gs, 2gs, ..], x2 <- last gs, y2 <- [gs-1, 2gs-1, ..]
    const int e = threadIdx.x;
    float num = 0.f, den = 0.f;
    for (int i = 0; i < gs; i++) {
      const int idx = e == 0 ? i : e == 1 ? i * gs : e == 2 ? P - gs + i : (i + 1) * gs - 1;
      const float a = (e & 1) ? ay[idx] : ax[idx];
      num += a * sc[idx];
      den += sc[idx];
    }
    out_boxes[4 * r + e] = num / den;
  }
}

// ---- streaming variant for large R ------------------------------------------------------------------------------
// The logits of one RoI are one contiguous block (P * h * w * 4 bytes, 28 224 for the 3x3 grid on 28x28 maps).  A persistent
// CTA walks RoIs blockIdx.x, blockIdx.x + gridDim.x, ...: one thread keeps the blocks of the next two RoIs under way with
// cp.async.bulk (completion on an mbarrier per stage), every warp takes its grid point's map out of shared memory into
// registers and reduces it exactly as above.  One block barrier per RoI: it publishes the points' results to the four
// voting threads and frees the stage for the RoI two ahead.  The one-CTA-per-RoI kernel above pays the full memory latency
// once per CTA with nothing else to do; here the copies never stop.
constexpr int kStreamStages = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
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

__global__ void __launch_bounds__(288, 3) grid_decode_stream_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                                  long R, int P, int gs, int h, int w, SubXY sub, float ratio,
                                                                  float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                                  uint32_t roi_bytes, uint32_t stage_bytes, unsigned wmagic) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ uint64_t full[kStreamStages];
  __shared__ float sc[kStreamStages][kMaxPoints], ax[kStreamStages][kMaxPoints], ay[kStreamStages][kMaxPoints];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = h * w;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStreamStages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long r, int s) {
    mbar_expect_tx(&full[s], roi_bytes);
    bulk_g2s(stage_mem + (size_t)s * stage_bytes, logits + r * (long)P * hw, roi_bytes, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < kStreamStages; s++)
      if ((long)blockIdx.x + (long)s * gridDim.x < R) issue((long)blockIdx.x + (long)s * gridDim.x, s);
  int s = 0;
  uint32_t parity = 0;
  float4 bx_next = (long)blockIdx.x < R ? __ldg(reinterpret_cast<const float4*>(boxes + 4 * (long)blockIdx.x))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long r = blockIdx.x; r < R; r += gridDim.x) {
    const float4 bx = bx_next;           // fetched one RoI ahead: its latency hides under this RoI's reduction
    if (r + gridDim.x < R) bx_next = __ldg(reinterpret_cast<const float4*>(boxes + 4 * (r + gridDim.x)));
    const float width = bx.z - bx.x, height = bx.w - bx.y;
    const float x1 = bx.x - ratio * (width / 2), y1 = bx.y - ratio * (height / 2);
    mbar_wait(&full[s], parity);
    const float* maps = reinterpret_cast<const float*>(stage_mem + (size_t)s * stage_bytes);
    for (int p = warp; p < P; p += nwarps) {
      float bs = -1.f;
      int bi = 0x7fffffff;
      point_argmax_regs<true>(reinterpret_cast<const float4*>(maps + (size_t)p * hw), hw >> 2, lane, bs, bi);
      if (lane == 0) {
        if (bi == 0x7fffffff) bi = 0;   // all-NaN map
        const int row = wmagic ? (int)(((unsigned)bi * wmagic) >> 16) : bi / w;
      const int xs = bi - row * w + sub.v[2 * p], ys = row + sub.v[2 * p + 1];
        sc[s][p] = bs;
        ax[s][p] = ((float)xs + 0.5f) / (float)(2 * w) * (1.f + ratio) * width + x1;      // inference.py:246
        ay[s][p] = ((float)ys + 0.5f) / (float)(2 * h) * (1.f + ratio) * height + y1;     // inference.py:247
        if (out_scores) out_scores[r * P + p] = bs;
      }
    }
    __syncthreads();     // results visible to the voting threads; every warp is done with stage s
    if (threadIdx.x == 0 && r + (long)kStreamStages * gridDim.x < R) issue(r + (long)kStreamStages * gridDim.x, s);
    if (threadIdx.x < 4) {
      const int e = threadIdx.x;      // boundary votes, inference.py:251-271 (as in grid_decode_kernel)
      float num = 0.f, den = 0.f;
      for (int i = 0; i < gs; i++) {
        const int idx = e == 0 ? i : e == 1 ? i * gs : e == 2 ? P - gs + i : (i + 1) * gs - 1;
        const float a = (e & 1) ? ay[s][idx] : ax[s][idx];
        num += a * sc[s][idx];
        den += sc[s][idx];
      }
      out_boxes[4 * r + e] = num / den;
    }
    if (++s == kStreamStages) {
      s = 0;
      parity ^= 1u;
    }
  }
}

// LevelMapper alone (poolers.py:29-40)
__global__ void level_map_kernel(const float* __restrict__ rois, long K, MapperView mp, long long* __restrict__ levels) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float* r = rois + 5 * i;
  levels[i] = fpn_level(r[1], r[2], r[3], r[4], mp);
}

}  // namespace cpm

using namespace cpm;

extern "C" int cpm_grid_decode(const float* d_logits, const float* d_boxes, int64_t R, int P, int h, int w,
                               const int32_t* sub_xy, float mapping_ratio, float* d_out_boxes, float* d_out_scores,
                               void* stream) {
  CPM_CHECK_ARG(R >= 0 && R < (1LL << 31), "R out of range");
  CPM_CHECK_ARG(P >= 1 && P <= kMaxPoints, "P must be in [1,%d]", kMaxPoints);
  int gs = 1;
  while (gs * gs < P) gs++;
  CPM_CHECK_ARG(gs * gs == P, "P must be a square number (grid_size^2)");
  CPM_CHECK_ARG(h >= 1 && w >= 1, "empty heat-map");
  CPM_CHECK_ARG(sub_xy != nullptr, "sub_xy is NULL");
  if (R == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_logits, "logits")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_boxes, "boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_out_boxes, "out_boxes")) != CPM_OK) return rc;
  CPM_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0 && ((uintptr_t)d_logits & 15) == 0, "logits/boxes must be 16-byte aligned");
  SubXY sub;
  for (int i = 0; i < 2 * P; i++) sub.v[i] = sub_xy[i];
  for (int i = 2 * P; i < 2 * kMaxPoints; i++) sub.v[i] = 0;
  int warps = P < 8 ? P : 8;
  if (P == 9) warps = 9;
  // i / w == (i * wmagic) >> 16 for every index of the map (checked here; 0 = divide in the kernel)
  unsigned wmagic = (65536u + (unsigned)w - 1) / (unsigned)w;
  if ((long)h * w > 4096) wmagic = 0;
  for (int i = 0; wmagic && i < h * w; i++)
    if ((((unsigned)i * wmagic) >> 16) != (unsigned)(i / w)) wmagic = 0;
  // large batches stream the RoIs through persistent CTAs (bulk copies, two RoIs ahead); small ones are latency-bound
  // anyway and keep one CTA per RoI
  const int hw = h * w;
  const size_t roi_bytes = (size_t)P * hw * sizeof(float);
  const size_t stage_bytes = (roi_bytes + 127) & ~(size_t)127;
  static const long stream_min = [] { const char* e = getenv("CPM_DECODE_STREAM_MIN"); return e ? atol(e) : 2048L; }();
  if (R >= stream_min && (hw & 3) == 0 && hw <= 4 * 32 * kRegVec && stage_bytes * kStreamStages <= 72 * 1024) {
    const size_t smem = stage_bytes * kStreamStages;
    CPM_CHECK_CUDA(cudaFuncSetAttribute(grid_decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long ctas = R < 148L * 3 ? R : 148L * 3;
    grid_decode_stream_kernel<<<(unsigned)ctas, 32 * warps, smem, (cudaStream_t)stream>>>(
        d_logits, d_boxes, R, P, gs, h, w, sub, mapping_ratio, d_out_boxes, d_out_scores, (uint32_t)roi_bytes,
        (uint32_t)stage_bytes, wmagic);
    CPM_CHECK_LAUNCH();
    return CPM_OK;
  }
  grid_decode_kernel<<<(unsigned)R, 32 * warps, 0, (cudaStream_t)stream>>>(d_logits, d_boxes, P, gs, h, w, sub,
                                                                          mapping_ratio, d_out_boxes, d_out_scores, wmagic);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}

extern "C" int cpm_level_map(const float* d_rois, int64_t K, const cpm_level_mapper_t* mapper, int64_t* d_levels,
                             void* stream) {
  CPM_CHECK_ARG(K >= 0, "K < 0");
  CPM_CHECK_ARG(mapper != nullptr, "mapper is NULL");
  if (K == 0) return CPM_OK;
  int rc;
  if ((rc = check_device_ptr(d_rois, "rois")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_levels, "levels")) != CPM_OK) return rc;
  level_map_kernel<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_rois, K, make_view(mapper),
                                                                                (long long*)d_levels);
  CPM_CHECK_LAUNCH();
  return CPM_OK;
}
