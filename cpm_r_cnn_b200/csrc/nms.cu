// Hard NMS for sm_100a: class-agnostic, label-gated (ml_nms) and batched-by-segment, entirely on the device.
//
// Replaces  pet.lib.ops.nms = torchvision.ops.nms                         (pet/lib/ops/nms.py:2,10)
//           _C.ml_nms  (pet/lib/ops/csrc/NMS/ml_nms.h:16-39, ml_nms.cu:82-146: sort, IoU bitmask kernel :29-78,
//                       mask D2H :117, serial host sweep :127-140)
//           and the per-image / per-level / per-class Python loops around them (rpn/inference.py:102-113,
//           grid_cascade_rcnn/inference.py:91-97, fast_rcnn/inference.py:105-164).
//
// Pipeline (all asynchronous on the caller's stream, no host round trip):
//   1. nms_build_keys      key = (segment << 32) | descending-orderable(score), value = input index
//   2. radix sort          stable => ties in score keep ascending input index (torchvision sorts stable; ml_nms.cu:92
//                          leaves tie order unspecified)
//   3. nms_gather          boxes into sorted order + segment-head flags -> compacted list of segment starts
//   4. nms_sweep           persistent CTAs, one segment at a time: the segment's boxes live in shared memory; 64 boxes
//                          per phase: a 64x64 IoU bit-matrix of the phase (shared-memory tile), a register-resident
//                          serial resolve of that matrix, then every surviving later box tests itself against the
//                          phase's kept boxes.  Only kept boxes ever act as suppressors, so the N x N/64 mask of the
//                          reference (and its D2H copy) never exists.
//   5. compaction          kept input indices in sorted order (+ a second sort by score for ml_nms, whose result is
//                          ordered by score across labels, ml_nms.cu:143-145).
// The IoU test is the reference's expression  inter / (Sa + Sb - inter) > thr  with IEEE division; `flavor` pins how the
// union is rounded (see include/cpm_ops.h), which is what makes keep indices bit-exact against each reference build.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace cpm {

int check_device_ptr(const void* p, const char* what);

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, const float thr, const int flavor) {
  // a = higher-scored (row) box, b = candidate (column) box; ml_nms.cu:19-25
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float aw = __fsub_rn(a.z, a.x), ah = __fsub_rn(a.w, a.y);
  const float bw = __fsub_rn(b.z, b.x), bh = __fsub_rn(b.w, b.y);
  float uni;
  if (flavor == CPM_IOU_TV_CUDA) {
    uni = __fsub_rn(__fmaf_rn(bw, bh, __fmul_rn(aw, ah)), inter);
  } else if (flavor == CPM_IOU_ML_CUDA) {
    uni = __fsub_rn(__fmaf_rn(aw, ah, __fmul_rn(bw, bh)), inter);
  } else {
    uni = __fsub_rn(__fadd_rn(__fmul_rn(aw, ah), __fmul_rn(bw, bh)), inter);
  }
  // The reference's decision is  fl(inter / uni) > thr.  Away from the threshold it follows from one multiplication: with
  // t = fl(thr * uni) (relative error 2^-24) and a guard band of 2^-20, inter outside [t (1 - 2^-20), t (1 + 2^-20)] decides
  // the rounded quotient's comparison as well -- branch-free for all but the pairs inside the band (and degenerate unions),
  // which take the exact IEEE division.  (A warp tests 32 rows against one column: with the division on the common path,
  // one overlapping lane made all 32 pay for it.)
  const float t = __fmul_rn(thr, uni);
  const float t1 = __fmul_rn(t, 1.00000095367431640625f), t0 = __fmul_rn(t, 0.99999904632568359375f);
  const bool pos = uni > 0.f && thr >= 0.f;
  if (pos && inter > t1) return true;
  if (pos && inter < t0) return false;
  if (inter == 0.f && thr >= 0.f) return false;   // 0/u and 0/0 (NaN) both compare false
  return __fdiv_rn(inter, uni) > thr;
}

__device__ __forceinline__ unsigned desc_key(float s) {
  // ascending unsigned order of the result == descending order of s; NaN first (torch.sort), -0 == +0
  if (s != s) return 0u;
  s += 0.0f;
  unsigned u = __float_as_uint(s);
  u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
  return ~u;
}

// trash >= 0: segment ids outside [0, trash) -- negative ones included -- are redirected to segment `trash`, which the sweep
// drops as a whole (a stray id can then neither write seg_counts out of bounds nor split a real segment)
template <typename SegT>
__global__ void nms_build_keys(const float* __restrict__ scores, const SegT* __restrict__ segs, int N, long long trash,
                               unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned long long seg = 0ull;
  if (segs) {
    const long long v = (long long)segs[i];
    seg = (trash >= 0 && (v < 0 || v >= trash)) ? (unsigned long long)trash : (unsigned long long)(unsigned)v;
  }
  keys[i] = (seg << 32) | desc_key(scores[i]);
  vals[i] = i;
}

__global__ void nms_gather(const float* __restrict__ boxes, const unsigned long long* __restrict__ keys,
                           const int* __restrict__ vals, int N, float4* __restrict__ sboxes,
                           unsigned char* __restrict__ heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  sboxes[i] = __ldg(reinterpret_cast<const float4*>(boxes) + vals[i]);
  heads[i] = i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32);
}

constexpr int kSweepCap = 2048;         // boxes of a segment kept in shared memory (larger segments read global/L2)
constexpr int kSweepMaxSeg = 65536;     // suppression bits of one segment live in shared memory

// WIDE = false: 256 threads, the phase's 64 rows resolved from registers by one thread -- many resident CTAs, for many
//               segments (detection: image x class).
// WIDE = true : 1024 threads on the strip test, rows resolved by one warp with shuffles -- for few large segments (RPN:
//               image x level with ~1000 boxes each), where one CTA per segment leaves most of the machine idle.
template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 1024 : 256, 1) nms_sweep(const float4* __restrict__ sboxes, const int* __restrict__ starts,
                                                            const int* __restrict__ d_nseg, int N,
                                                            const unsigned long long* __restrict__ keys, float thr,
                                                            long long topk, int flavor, unsigned char* __restrict__ flags,
                                                            long long* __restrict__ seg_counts,
                                                            unsigned long long* __restrict__ total_kept,
                                                            unsigned* __restrict__ rem_g, long long trash, int skip_upto) {
  __shared__ float4 sb[kSweepCap];
  __shared__ unsigned rem_s[kSweepMaxSeg / 32];
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ int s_done;
  const int tid = threadIdx.x;
  constexpr int kSweepThreads = WIDE ? 1024 : 256;
  const int nseg = *d_nseg;
  for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
    const int beg = starts[s];
    const int end = s + 1 < nseg ? starts[s + 1] : N;
    const int n = end - beg;
    if (n <= skip_upto) continue;         // taken by the bitmask path (nms_mask + nms_sweep_mask)
    if (trash >= 0 && (long long)(keys[beg] >> 32) >= trash) {      // boxes with an out-of-range segment id: all dropped
      for (int j = tid; j < n; j += kSweepThreads) flags[beg + j] = 0;
      continue;
    }
    const bool in_smem = n <= kSweepCap;
    const float4* gb = sboxes + beg;
    if (in_smem)
      for (int j = tid; j < n; j += kSweepThreads) sb[j] = gb[j];
    // suppression bits: shared memory up to 65 536 boxes, a private slice of the workspace beyond (two big segments start
    // more than 65 536 boxes apart, so beg / 32 + 2 * (beg / 65536) never lets their word ranges touch)
    const bool big = n > kSweepMaxSeg;
    unsigned* const remg = rem_g + (beg >> 5) + 2 * (beg >> 16);
    auto sweep_segment = [&](unsigned* rem) {
    for (int j = tid; j < (n + 31) / 32; j += kSweepThreads) rem[j] = 0u;
    __syncthreads();
    long long kept_total = 0;
    for (int base = 0; base < n; base += 64) {
      const int m = min(64, n - base);
      // ---- 64x64 IoU bit-matrix of this phase: thread t -> row t/4, 16 columns ----
      if (tid < 64) diag[tid] = 0ull;
      __syncthreads();
      if (tid < 256) {
        const int i = tid >> 2, j0 = (tid & 3) * 16;
        if (i < m && !((rem[(base + i) >> 5] >> ((base + i) & 31)) & 1u)) {
          const float4 a = in_smem ? sb[base + i] : gb[base + i];
          unsigned long long bits = 0ull;
          for (int j = max(j0, i + 1); j < min(j0 + 16, m); j++) {
            const float4 b = in_smem ? sb[base + j] : gb[base + j];
            if (iou_gt(a, b, thr, flavor)) bits |= 1ull << j;
          }
          if (bits) atomicOr(&diag[i], bits);
        }
      }
      __syncthreads();
      // ---- serial resolve (ml_nms.cu:127-140 restricted to the phase), rows register-resident ----
      if (WIDE ? tid < 32 : tid == 0) {
        const unsigned lo = rem[base >> 5], hi = (base + 32 < n) ? rem[(base >> 5) + 1] : 0u;
        unsigned long long alive = ~(((unsigned long long)hi << 32) | lo);
        if (m < 64) alive &= (1ull << m) - 1;
        unsigned long long kept = 0ull;
        if (WIDE) {       // lane l holds rows l and l + 32; the row of every kept box is broadcast (alive / kept are uniform)
          const unsigned long long r0 = diag[tid & 31], r1 = diag[(tid & 31) + 32];
#pragma unroll
          for (int i = 0; i < 64; i++) {
            if ((alive >> i) & 1ull) {
              kept |= 1ull << i;
              alive &= ~__shfl_sync(0xffffffffu, i < 32 ? r0 : r1, i & 31);
            }
          }
        } else {
          unsigned long long row[64];
#pragma unroll
          for (int i = 0; i < 64; i++) row[i] = diag[i];
#pragma unroll
          for (int i = 0; i < 64; i++) {
            if ((alive >> i) & 1ull) {
              kept |= 1ull << i;
              alive &= ~row[i];
            }
          }
        }
        int done = 0;
        if (topk > 0 && kept_total + __popcll(kept) >= topk) {   // ml_nms.cu:134: stop after topk keeps
          long long room = topk - kept_total;
          unsigned long long k2 = 0ull, k = kept;
          while (room > 0 && k) {
            k2 |= k & (~k + 1);
            k &= k - 1;
            room--;
          }
          kept = k2;
          done = 1;
        }
        if (tid == 0) {
          s_kept = kept;
          s_done = done;
        }
      }
      __syncthreads();
      const unsigned long long kept = s_kept;
      const int done = s_done;
      if (tid < m) flags[beg + base + tid] = (unsigned char)((kept >> tid) & 1ull);
      kept_total += __popcll(kept);
      if (done) {
        for (int j = base + 64 + tid; j < n; j += kSweepThreads) flags[beg + j] = 0;
        break;
      }
      // ---- every surviving later box against the kept boxes of this phase ----
      if (kept) {
        for (int j = base + 64 + tid; j < n; j += kSweepThreads) {
          if ((rem[j >> 5] >> (j & 31)) & 1u) continue;
          const float4 b = in_smem ? sb[j] : gb[j];
          unsigned long long k = kept;
          while (k) {
            const int i = __ffsll((long long)k) - 1;
            k &= k - 1;
            const float4 a = in_smem ? sb[base + i] : gb[base + i];
            if (iou_gt(a, b, thr, flavor)) {
              atomicOr(&rem[j >> 5], 1u << (j & 31));
              break;
            }
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (seg_counts) seg_counts[keys[beg] >> 32] = kept_total;
      if (total_kept) atomicAdd(total_kept, (unsigned long long)kept_total);
    }
    __syncthreads();
    };
    if (big) sweep_segment(remg);
    else sweep_segment(rem_s);
  }
}

// ---- bitmask path: few large segments (RPN: image x level, ~1000 boxes each) ---------------------------------------
// The phased sweep above keeps one segment on one SM; with 80 segments more than a third of the machine idles.  Here the
// IoU tests of a segment are spread over the grid: nms_mask writes, per sorted box, one 64-bit word per 64-box column
// tile to its right (shared-memory box tiles; the reference's mask kernel, ml_nms.cu:29-78, without the D2H copy), and
// nms_sweep_mask resolves a segment with ONE warp (ml_nms.cu:127-140 on the device).
constexpr int kMaskWords = 32;                     // row stride: segments up to 2048 boxes
constexpr int kMaskMaxSeg = 64 * kMaskWords;
constexpr long long kMaskMaxN = 131072;            // 256 bytes of mask per box: 32 MB of workspace at most

__global__ void __launch_bounds__(256) nms_mask(const float4* __restrict__ sboxes, const int* __restrict__ starts,
                                                 const int* __restrict__ d_nseg, int N,
                                                 const unsigned long long* __restrict__ keys, float thr, int flavor,
                                                 long long trash, unsigned long long* __restrict__ mask) {
  __shared__ float4 cb[4][64];
  const int s = blockIdx.x;
  const int nseg = *d_nseg;
  if (s >= nseg) return;
  const int beg = starts[s];
  const int n = (s + 1 < nseg ? starts[s + 1] : N) - beg;
  const int rt = blockIdx.y;
  if (n > kMaskMaxSeg || rt * 64 >= n) return;
  if (trash >= 0 && (long long)(keys[beg] >> 32) >= trash) return;
  const int r = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int row = rt * 64 + r;
  const float4 a = row < n ? sboxes[beg + row] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int nt = (n + 63) >> 6;
  for (int jt = rt + g; jt < nt; jt += 4) {
    const int col = jt * 64 + r;
    cb[g][r] = col < n ? sboxes[beg + col] : make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
    unsigned long long bits = 0ull;
    const int j0 = jt == rt ? r + 1 : 0, j1 = min(64, n - jt * 64);
    if (row < n)
      for (int j = j0; j < j1; j++)
        if (iou_gt(a, cb[g][j], thr, flavor)) bits |= 1ull << j;
    if (row < n) mask[(size_t)(beg + row) * kMaskWords + jt] = bits;
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
  }
}

// One warp (one CTA) per segment.  The mask rows of a 64-box phase -- the words from the phase's own column tile to the
// right end -- are staged in shared memory with cp.async one phase ahead; the phase is resolved on registers (every lane
// holds the 64 diagonal words; `alive` and `kept` are warp-uniform, no shuffles in the chain), then lane w ORs the rows of
// the kept boxes into word w of the suppressed bitmap.
__global__ void __launch_bounds__(32) nms_sweep_mask(const int* __restrict__ starts, const int* __restrict__ d_nseg, int N,
                                                      const unsigned long long* __restrict__ keys, long long topk,
                                                      long long trash, const unsigned long long* __restrict__ mask,
                                                      unsigned char* __restrict__ flags, long long* __restrict__ seg_counts,
                                                      unsigned long long* __restrict__ total_kept) {
  __shared__ __align__(16) unsigned long long sm[2][64][kMaskWords];     // 32 KB
  const int lane = threadIdx.x;
  const int s = blockIdx.x;
  const int nseg = *d_nseg;
  if (s >= nseg) return;
  const int beg = starts[s];
  const int n = (s + 1 < nseg ? starts[s + 1] : N) - beg;
  if (n > kMaskMaxSeg) return;
  if (trash >= 0 && (long long)(keys[beg] >> 32) >= trash) {
    for (int j = lane; j < n; j += 32) flags[beg + j] = 0;
    return;
  }
  const int nt = (n + 63) >> 6;
  const unsigned long long* mrow = mask + (size_t)beg * kMaskWords;
  auto stage = [&](int ph, int buf) {
    const int base = ph * 64, m = min(64, n - base);
    const int c0 = ph >> 1, c1 = (nt + 1) >> 1;            // 16-byte chunks (word pairs) holding words ph .. nt-1
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = lane + 32 * h;
      if (r < m) {
        const unsigned long long* g = mrow + (size_t)(base + r) * kMaskWords;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(&sm[buf][r][0]);
        for (int c = c0; c < c1; c++)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * c), "l"(g + 2 * c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  unsigned long long remv = 0ull;                  // word `lane` of the segment's suppressed bitmap
  long long kept_total = 0;
  stage(0, 0);
  for (int ph = 0; ph < nt; ph++) {
    const int base = ph * 64, buf = ph & 1;
    const int m = min(64, n - base);
    if (ph + 1 < nt) {
      stage(ph + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    unsigned long long row[64];
#pragma unroll
    for (int i = 0; i < 64; i++) row[i] = sm[buf][i][ph];          // rows >= m are never consulted (alive is masked)
    unsigned long long alive = ~__shfl_sync(0xffffffffu, remv, ph);
    if (m < 64) alive &= (1ull << m) - 1;
    unsigned long long kept = 0ull;
#pragma unroll
    for (int i = 0; i < 64; i++) {
      if ((alive >> i) & 1ull) {
        kept |= 1ull << i;
        alive &= ~row[i];
      }
    }
    bool done = false;
    if (topk > 0 && kept_total + __popcll(kept) >= topk) {   // ml_nms.cu:134: stop after topk keeps
      long long room = topk - kept_total;
      unsigned long long k2 = 0ull, k = kept;
      while (room > 0 && k) {
        k2 |= k & (~k + 1);
        k &= k - 1;
        room--;
      }
      kept = k2;
      done = true;
    }
    if (lane < m) flags[beg + base + lane] = (unsigned char)((kept >> lane) & 1ull);
    if (lane + 32 < m) flags[beg + base + lane + 32] = (unsigned char)((kept >> (lane + 32)) & 1ull);
    kept_total += __popcll(kept);
    if (done) {
      for (int j = base + 64 + lane; j < n; j += 32) flags[beg + j] = 0;
      break;
    }
    if (lane > ph && lane < nt) {                  // rows of the kept boxes -> word `lane` (kept is warp-uniform)
      unsigned long long o = 0ull;
#pragma unroll
      for (int i = 0; i < 64; i++)
        if ((kept >> i) & 1ull) o |= sm[buf][i][lane];
      remv |= o;
    }
    __syncwarp();                                  // the buffer is restaged two phases later
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (lane == 0) {
    if (seg_counts) seg_counts[keys[beg] >> 32] = kept_total;
    if (total_kept) atomicAdd(total_kept, (unsigned long long)kept_total);
  }
}

// ml_nms: kept boxes re-keyed by (descending score, ascending input index); dropped boxes sort to the end
__global__ void nms_rekey(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                          const unsigned char* __restrict__ flags, int N, unsigned long long* __restrict__ keys2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keys2[i] = flags[i] ? ((keys[i] << 32) | (unsigned)vals[i]) : ~0ull;
}

__global__ void nms_finish_ml(const int* __restrict__ vals2, const unsigned long long* __restrict__ total_kept,
                              long long topk, int N, long long* __restrict__ keep, long long* __restrict__ count) {
  long long c = (long long)*total_kept;
  if (topk > 0 && c > topk) c = topk;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *count = c;
  if (i < N && i < c) keep[i] = vals2[i];
}

struct ToI64 {
  __host__ __device__ long long operator()(const int& v) const { return (long long)v; }
};

// ---- workspace carving ----
struct NmsWs {
  size_t keys_a, keys_b, vals_a, vals_b, sboxes, flags, heads, starts, scalars, rem, mask, cub, total;
  size_t cub_bytes;
};

static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

static NmsWs nms_layout(int64_t N) {
  NmsWs w;
  const size_t n = (size_t)(N > 0 ? N : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  w.keys_a = take(n * 8);
  w.keys_b = take(n * 8);
  w.vals_a = take(n * 4);
  w.vals_b = take(n * 4);
  w.sboxes = take(n * 16);
  w.flags = take(n);
  w.heads = take(n);
  w.starts = take((n + 1) * 4);
  w.scalars = take(64);
  w.rem = take((n / 32 + 2 * (n / 65536) + 8) * 4);     // suppression bits of segments above 65 536 boxes
  w.mask = take((long long)n <= kMaskMaxN ? n * kMaskWords * 8 : 0);   // IoU bitmask of the few-large-segments path
  // cub temp storage: asked from cub when a device is present, else a documented upper bound
  size_t t1 = 0, t2 = 0, t3 = 0;
  cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr,
                                                   (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr,
                                                   (int)n, 0, 64, (cudaStream_t)0);
  cub::CountingInputIterator<int> cnt(0);
  cudaError_t e2 = cub::DeviceSelect::Flagged(nullptr, t2, cnt, (const unsigned char*)nullptr, (int*)nullptr,
                                              (int*)nullptr, (int)n, (cudaStream_t)0);
  cub::TransformInputIterator<long long, ToI64, const int*> tin((const int*)nullptr, ToI64());
  cudaError_t e3 = cub::DeviceSelect::Flagged(nullptr, t3, tin, (const unsigned char*)nullptr, (long long*)nullptr,
                                              (long long*)nullptr, (int)n, (cudaStream_t)0);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    cudaGetLastError();
    t1 = n * 16 + (8u << 20);
    t2 = t3 = n * 8 + (1u << 20);
  }
  w.cub_bytes = t1 > t2 ? t1 : t2;
  if (t3 > w.cub_bytes) w.cub_bytes = t3;
  w.cub = take(w.cub_bytes);
  w.total = off;
  return w;
}

static int bits_for(int64_t v) {
  int b = 0;
  while (b < 32 && ((int64_t)1 << b) < v) b++;
  return b;
}

// mode 0: single segment; 1: int32 segments (batched); 2: int64 labels (ml_nms)
static int run_nms(const float* d_boxes, const float* d_scores, const void* d_segs, int mode, int64_t N,
                   int64_t num_segments, float thr, int64_t topk, int flavor, int64_t* d_keep, int64_t* d_seg_counts,
                   int64_t* d_count, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  CPM_CHECK_ARG(N >= 0 && N < (1LL << 31), "N out of range");
  CPM_CHECK_ARG(flavor >= CPM_IOU_PLAIN && flavor <= CPM_IOU_ML_CUDA, "unknown iou flavor %d", flavor);
  int rc;
  if ((rc = check_device_ptr(d_count, "count")) != CPM_OK) return rc;
  if (d_seg_counts && num_segments > 0)
    CPM_CHECK_CUDA(cudaMemsetAsync(d_seg_counts, 0, (size_t)num_segments * sizeof(int64_t), st));
  if (N == 0) {
    CPM_CHECK_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    return CPM_OK;
  }
  if ((rc = check_device_ptr(d_boxes, "boxes")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_scores, "scores")) != CPM_OK) return rc;
  if ((rc = check_device_ptr(d_keep, "keep")) != CPM_OK) return rc;
  CPM_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0, "boxes must be 16-byte aligned");
  const NmsWs w = nms_layout(N);
  if (d_ws == nullptr || ws_bytes < w.total) {
    set_error("workspace too small: %zu < %zu bytes", ws_bytes, w.total);
    return CPM_ERR_WORKSPACE;
  }
  if ((rc = check_device_ptr(d_ws, "workspace")) != CPM_OK) return rc;
  char* base = (char*)d_ws;
  auto* keys_a = (unsigned long long*)(base + w.keys_a);
  auto* keys_b = (unsigned long long*)(base + w.keys_b);
  int* vals_a = (int*)(base + w.vals_a);
  int* vals_b = (int*)(base + w.vals_b);
  float4* sboxes = (float4*)(base + w.sboxes);
  unsigned char* flags = (unsigned char*)(base + w.flags);
  unsigned char* heads = (unsigned char*)(base + w.heads);
  int* starts = (int*)(base + w.starts);
  int* d_nseg = (int*)(base + w.scalars);
  unsigned* rem_g = (unsigned*)(base + w.rem);
  unsigned long long* d_total = (unsigned long long*)(base + w.scalars + 16);
  void* cub_tmp = base + w.cub;
  size_t cub_bytes = w.cub_bytes;

  const int n = (int)N;
  const int tb = 256, gb = (n + tb - 1) / tb;
  CPM_CHECK_CUDA(cudaMemsetAsync(base + w.scalars, 0, 64, st));
  const long long trash = mode == 1 ? (long long)num_segments : -1LL;
  if (mode == 2)
    nms_build_keys<long long><<<gb, tb, 0, st>>>(d_scores, (const long long*)d_segs, n, -1LL, keys_a, vals_a);
  else
    nms_build_keys<int><<<gb, tb, 0, st>>>(d_scores, mode == 1 ? (const int*)d_segs : nullptr, n, trash, keys_a, vals_a);
  CPM_CHECK_LAUNCH();
  const int end_bit = mode == 0 ? 32 : (mode == 1 ? 32 + bits_for(num_segments + 1) : 64);
  CPM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, vals_a, vals_b, n, 0, end_bit, st));
  count_launch(3);
  nms_gather<<<gb, tb, 0, st>>>(d_boxes, keys_b, vals_b, n, sboxes, heads);
  CPM_CHECK_LAUNCH();
  cub::CountingInputIterator<int> cnt(0);
  cub_bytes = w.cub_bytes;
  CPM_CHECK_CUDA(cub::DeviceSelect::Flagged(cub_tmp, cub_bytes, cnt, heads, starts, d_nseg, n, st));
  count_launch(2);
  const int64_t segs = mode == 1 ? (num_segments > 0 ? num_segments : 1) : 1;
  const bool wide = mode == 0 || (n / segs >= 256 && segs < 148 * 4);
  const long long sweep_topk = mode == 2 ? 0LL : (long long)topk;
  long long* seg_counts = mode == 1 ? (long long*)d_seg_counts : nullptr;
  // few large segments (or one segment of at most 2048 boxes): the IoU bitmask is built by the whole grid and every
  // segment of at most 2048 boxes is resolved by one warp; larger segments of the same call stay with the phased sweep
  const bool masked = mode != 2 && wide && (long long)n <= kMaskMaxN && (mode == 1 || n <= kMaskMaxSeg);
  int skip_upto = 0;
  if (masked) {
    unsigned long long* mask = (unsigned long long*)(base + w.mask);
    nms_mask<<<dim3((unsigned)segs, kMaskWords), 256, 0, st>>>(sboxes, starts, d_nseg, n, keys_b, thr, flavor, trash, mask);
    CPM_CHECK_LAUNCH();
    nms_sweep_mask<<<(unsigned)segs, 32, 0, st>>>(starts, d_nseg, n, keys_b, sweep_topk, trash, mask, flags, seg_counts,
                                                  d_total);
    CPM_CHECK_LAUNCH();
    skip_upto = kMaskMaxSeg;
  }
  if (wide)
    nms_sweep<true><<<mode == 0 ? 1 : 148 * 2, 1024, 0, st>>>(sboxes, starts, d_nseg, n, keys_b, thr, sweep_topk, flavor, flags,
                                                              seg_counts, d_total, rem_g, trash, skip_upto);
  else
    nms_sweep<false><<<148 * 4, 256, 0, st>>>(sboxes, starts, d_nseg, n, keys_b, thr, sweep_topk, flavor, flags, seg_counts,
                                              d_total, rem_g, trash, skip_upto);
  CPM_CHECK_LAUNCH();
  if (mode == 2) {
    nms_rekey<<<gb, tb, 0, st>>>(keys_b, vals_b, flags, n, keys_a);
    CPM_CHECK_LAUNCH();
    cub_bytes = w.cub_bytes;
    CPM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, vals_b, vals_a, n, 0, 64, st));
    count_launch(3);
    nms_finish_ml<<<gb, tb, 0, st>>>(vals_a, d_total, (long long)topk, n, (long long*)d_keep, (long long*)d_count);
    CPM_CHECK_LAUNCH();
  } else {
    cub::TransformInputIterator<long long, ToI64, const int*> tin(vals_b, ToI64());
    cub_bytes = w.cub_bytes;
    CPM_CHECK_CUDA(cub::DeviceSelect::Flagged(cub_tmp, cub_bytes, tin, flags, (long long*)d_keep, (long long*)d_count, n, st));
    count_launch(2);
  }
  return CPM_OK;
}

}  // namespace cpm

using namespace cpm;

extern "C" size_t cpm_nms_workspace_bytes(int64_t N) { return nms_layout(N).total; }

extern "C" size_t cpm_nms_batched_workspace_bytes(int64_t N, int64_t num_segments) {
  (void)num_segments;
  return nms_layout(N).total;
}

extern "C" int cpm_nms(const float* d_boxes, const float* d_scores, const int64_t* d_labels, int64_t N, float iou_threshold,
                       int64_t topk, int iou_flavor, int64_t* d_keep, int64_t* d_count, void* d_workspace,
                       size_t workspace_bytes, void* stream) {
  if (d_labels != nullptr && N > 0) {
    int rc = check_device_ptr(d_labels, "labels");
    if (rc != CPM_OK) return rc;
  }
  return run_nms(d_boxes, d_scores, d_labels, d_labels ? 2 : 0, N, 1, iou_threshold, topk, iou_flavor, d_keep, nullptr,
                 d_count, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cpm_nms_batched(const float* d_boxes, const float* d_scores, const int32_t* d_segments, int64_t N,
                               int64_t num_segments, float iou_threshold, int64_t topk_per_segment, int iou_flavor,
                               int64_t* d_keep, int64_t* d_seg_counts, int64_t* d_count, void* d_workspace,
                               size_t workspace_bytes, void* stream) {
  CPM_CHECK_ARG(num_segments >= 1 && num_segments < (1LL << 31), "num_segments out of range");
  if (N > 0) {
    int rc = check_device_ptr(d_segments, "segments");
    if (rc != CPM_OK) return rc;
  }
  return run_nms(d_boxes, d_scores, d_segments, 1, N, num_segments, iou_threshold, topk_per_segment, iou_flavor, d_keep,
                 d_seg_counts, d_count, d_workspace, workspace_bytes, (cudaStream_t)stream);
}
