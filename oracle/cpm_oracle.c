/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's detection-head op
 * layer.  Nothing under cpm_r_cnn_b200/ may include, link or call this file; it exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path.
 *
 * Parity pin: every function below is checked (tests/test_oracle_cpu.py) against the reference's own
 * compiled sources (oracle/_ref, built by oracle/build_ref.py from /root/reference, unmodified) or
 * against torchvision 0.26 (the un-vendored third-party `nms` the reference imports at
 * pet/lib/ops/nms.py:2), and against the committed fixtures under tests/golden/.
 *
 * Plain C, scalar, single-threaded, compiled with -ffp-contract=off so that every rounding step is
 * the one written here.  Reference file:line is cited on each function.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * RoIAlign.  Follows pet/lib/ops/csrc/ROIAlign/ROIAlign_cuda.cu:178-256 (forward kernel),
 * :36-86 (bilinear_interpolate), :14-33 (nearest_interpolate); the CPU twin
 * ROIAlign_cpu.cpp:73-166,169-294 computes the same expression with taps hoisted out of the channel loop.
 * `strict_aligned_check` reproduces the CPU-only assertion ROIAlign_cpu.cpp:202-205 (returns -1).
 * ---------------------------------------------------------------------------------------------- */
#define ORC_DEFINE_ROI_ALIGN(T, SUF)                                                                  \
  typedef struct { int valid; int yl, yh, xl, xh; T w1, w2, w3, w4; } orc_tap_##SUF;                 \
                                                                                                      \
  /* ROIAlign_cuda.cu:113-171 / ROIAlign_cpu.cpp:321-379 */                                           \
  static void orc_bilinear_tap_##SUF(int H, int W, T y, T x, orc_tap_##SUF* t) {                      \
    if (y < (T)-1.0 || y > (T)H || x < (T)-1.0 || x > (T)W) {                                         \
      t->valid = 0; t->yl = t->yh = t->xl = t->xh = -1; t->w1 = t->w2 = t->w3 = t->w4 = 0; return;    \
    }                                                                                                 \
    if (y <= 0) y = 0;                                                                                \
    if (x <= 0) x = 0;                                                                                \
    int yl = (int)y, xl = (int)x, yh, xh;                                                             \
    if (yl >= H - 1) { yh = yl = H - 1; y = (T)yl; } else { yh = yl + 1; }                            \
    if (xl >= W - 1) { xh = xl = W - 1; x = (T)xl; } else { xh = xl + 1; }                            \
    T ly = y - yl, lx = x - xl, hy = (T)1. - ly, hx = (T)1. - lx;                                     \
    t->valid = 1; t->yl = yl; t->yh = yh; t->xl = xl; t->xh = xh;                                     \
    t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;                               \
  }                                                                                                   \
                                                                                                      \
  /* ROIAlign_cuda.cu:14-33, :89-110 */                                                               \
  static int orc_nearest_tap_##SUF(int H, int W, T y, T x, int* yl, int* xl) {                        \
    if (y < (T)-0.5 || y >= (T)H - (T)0.5 || x < (T)-0.5 || x >= (T)W - (T)0.5) {                     \
      *yl = *xl = -1; return 0;                                                                       \
    }                                                                                                 \
    *xl = (int)round((double)x); *yl = (int)round((double)y); return 1;                               \
  }                                                                                                   \
                                                                                                      \
  typedef struct { T start_w, start_h, bin_w, bin_h; int gh, gw, b; int bad; } orc_geo_##SUF;         \
                                                                                                      \
  /* ROIAlign_cuda.cu:199-230 */                                                                      \
  static orc_geo_##SUF orc_geometry_##SUF(const T* roi, T scale, int PH, int PW, int sr, int aligned) { \
    orc_geo_##SUF g;                                                                                  \
    g.b = (int)roi[0];                                                                                \
    T off = aligned ? (T)0.5 : (T)0.0;                                                                \
    g.start_w = roi[1] * scale - off; g.start_h = roi[2] * scale - off;                               \
    T end_w = roi[3] * scale - off, end_h = roi[4] * scale - off;                                     \
    T rw = end_w - g.start_w, rh = end_h - g.start_h;                                                 \
    g.bad = aligned && !(rw >= 0 && rh >= 0);                                                         \
    if (!aligned) { rw = rw > (T)1. ? rw : (T)1.; rh = rh > (T)1. ? rh : (T)1.; }                     \
    g.bin_h = rh / (T)PH; g.bin_w = rw / (T)PW;                                                       \
    g.gh = sr > 0 ? sr : (int)ceil((double)(rh / PH));                                                \
    g.gw = sr > 0 ? sr : (int)ceil((double)(rw / PW));                                                \
    return g;                                                                                         \
  }                                                                                                   \
                                                                                                      \
  int orc_roi_align_fwd_##SUF(const T* feat, int B, int C, int H, int W, const T* rois, int K,        \
                              T scale, int PH, int PW, int sr, int aligned, int interp,               \
                              int strict_aligned_check, T* out) {                                     \
    (void)B;                                                                                          \
    for (int n = 0; n < K; n++) {                                                                     \
      orc_geo_##SUF g = orc_geometry_##SUF(rois + 5 * n, scale, PH, PW, sr, aligned);                 \
      if (g.bad && strict_aligned_check) return -1;                                                   \
      int cnt_i = g.gh * g.gw; if (cnt_i < 1) cnt_i = 1;                                              \
      const T count = (T)cnt_i;                                                                       \
      for (int c = 0; c < C; c++) {                                                                   \
        const T* f = feat + ((size_t)g.b * C + c) * H * W;                                            \
        for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {                             \
          T acc = 0;                                                                                  \
          for (int iy = 0; iy < g.gh; iy++) {                                                         \
            const T y = g.start_h + ph * g.bin_h + (T)(iy + .5f) * g.bin_h / (T)g.gh;                 \
            for (int ix = 0; ix < g.gw; ix++) {                                                       \
              const T x = g.start_w + pw * g.bin_w + (T)(ix + .5f) * g.bin_w / (T)g.gw;               \
              if (interp == 0) {                                                                      \
                orc_tap_##SUF t; orc_bilinear_tap_##SUF(H, W, y, x, &t);                              \
                if (t.valid)                                                                          \
                  acc += (t.w1 * f[t.yl * W + t.xl] + t.w2 * f[t.yl * W + t.xh] +                     \
                          t.w3 * f[t.yh * W + t.xl] + t.w4 * f[t.yh * W + t.xh]);                     \
                else acc += 0;                                                                        \
              } else {                                                                                \
                int yl, xl;                                                                           \
                if (orc_nearest_tap_##SUF(H, W, y, x, &yl, &xl)) acc += f[yl * W + xl];               \
              }                                                                                       \
            }                                                                                         \
          }                                                                                           \
          out[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc / count;                               \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
    return 0;                                                                                         \
  }                                                                                                   \
                                                                                                      \
  /* ROIAlign_cuda.cu:259-365 (scatter; serial here, so summation order = (n,c,ph,pw,iy,ix)),         \
     identical to ROIAlign_cpu.cpp:387-496.  grad_in (B,C,H,W) is zero-filled first (:451-452). */    \
  int orc_roi_align_bwd_##SUF(const T* grad_out, const T* rois, int K, T scale, int PH, int PW,       \
                              int B, int C, int H, int W, int sr, int aligned, int interp,            \
                              int strict_aligned_check, T* grad_in) {                                 \
    memset(grad_in, 0, sizeof(T) * (size_t)B * C * H * W);                                            \
    for (int n = 0; n < K; n++) {                                                                     \
      orc_geo_##SUF g = orc_geometry_##SUF(rois + 5 * n, scale, PH, PW, sr, aligned);                 \
      if (g.bad && strict_aligned_check) return -1;                                                   \
      const T count = (T)(g.gh * g.gw);                                                               \
      for (int c = 0; c < C; c++) {                                                                   \
        T* gi = grad_in + ((size_t)g.b * C + c) * H * W;                                              \
        for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {                             \
          const T go = grad_out[(((size_t)n * C + c) * PH + ph) * PW + pw];                           \
          for (int iy = 0; iy < g.gh; iy++) {                                                         \
            const T y = g.start_h + ph * g.bin_h + (T)(iy + .5f) * g.bin_h / (T)g.gh;                 \
            for (int ix = 0; ix < g.gw; ix++) {                                                       \
              const T x = g.start_w + pw * g.bin_w + (T)(ix + .5f) * g.bin_w / (T)g.gw;               \
              if (interp == 0) {                                                                      \
                orc_tap_##SUF t; orc_bilinear_tap_##SUF(H, W, y, x, &t);                              \
                T g1 = go * t.w1 / count, g2 = go * t.w2 / count;                                     \
                T g3 = go * t.w3 / count, g4 = go * t.w4 / count;                                     \
                if (t.xl >= 0 && t.xh >= 0 && t.yl >= 0 && t.yh >= 0) {                               \
                  gi[t.yl * W + t.xl] += g1; gi[t.yl * W + t.xh] += g2;                               \
                  gi[t.yh * W + t.xl] += g3; gi[t.yh * W + t.xh] += g4;                               \
                }                                                                                     \
              } else {                                                                                \
                int yl, xl;                                                                           \
                if (orc_nearest_tap_##SUF(H, W, y, x, &yl, &xl)) gi[yl * W + xl] += go / count;       \
              }                                                                                       \
            }                                                                                         \
          }                                                                                           \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
    return 0;                                                                                         \
  }                                                                                                   \
                                                                                                      \
  /* Number of distinct (y,x) pixels of one (H,W) level read by any bilinear tap of any RoI with      \
     batch index b_sel (-1 = all images share one mask per image; mask is B*H*W bytes, caller-zeroed). \
     Used for SURVEY.md 8(d)'s U (algorithmic forward bytes). */                                      \
  void orc_roi_align_touch_##SUF(const T* rois, int K, T scale, int PH, int PW, int sr, int aligned,  \
                                 int B, int H, int W, uint8_t* mask) {                                \
    for (int n = 0; n < K; n++) {                                                                     \
      orc_geo_##SUF g = orc_geometry_##SUF(rois + 5 * n, scale, PH, PW, sr, aligned);                 \
      if (g.b < 0 || g.b >= B) continue;                                                              \
      uint8_t* m = mask + (size_t)g.b * H * W;                                                        \
      for (int ph = 0; ph < PH; ph++) for (int iy = 0; iy < g.gh; iy++) {                             \
        const T y = g.start_h + ph * g.bin_h + (T)(iy + .5f) * g.bin_h / (T)g.gh;                     \
        for (int pw = 0; pw < PW; pw++) for (int ix = 0; ix < g.gw; ix++) {                           \
          const T x = g.start_w + pw * g.bin_w + (T)(ix + .5f) * g.bin_w / (T)g.gw;                   \
          orc_tap_##SUF t; orc_bilinear_tap_##SUF(H, W, y, x, &t);                                    \
          if (!t.valid) continue;                                                                     \
          m[t.yl * W + t.xl] = 1; m[t.yl * W + t.xh] = 1; m[t.yh * W + t.xl] = 1; m[t.yh * W + t.xh] = 1; \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
  }

ORC_DEFINE_ROI_ALIGN(float, f32)
ORC_DEFINE_ROI_ALIGN(double, f64)

/* ------------------------------------------------------------------------------------------------
 * FPN level mapper.  pet/rcnn/utils/poolers.py:29-40 with BoxList.area() in xyxy mode
 * (pet/utils/data/structures/bounding_box.py:306-310, TO_REMOVE = 1) and k_min/k_max from
 * poolers.py:86-88.  rois are (K,5) [img, x1, y1, x2, y2]; output = level index in [0, k_max-k_min].
 * `recip_div` = 1 restates torch's CUDA `tensor / python_scalar` (a * (1/b)), 0 = true division (CPU).
 * ---------------------------------------------------------------------------------------------- */
void orc_level_map(const float* rois, int K, float k_min, float k_max, float s0, float lvl0, float eps,
                   int recip_div, int64_t* levels) {
  const float inv_s0 = 1.0f / s0;
  for (int i = 0; i < K; i++) {
    const float* r = rois + 5 * i;
    float area = (r[3] - r[1] + 1.0f) * (r[4] - r[2] + 1.0f);
    float s = sqrtf(area);
    float q = recip_div ? s * inv_s0 : s / s0;
    float t = floorf(lvl0 + log2f(q + eps));
    if (t < k_min) t = k_min;  /* torch.clamp: NaN propagates; not reachable for finite boxes */
    if (t > k_max) t = k_max;
    levels[i] = (int64_t)t - (int64_t)k_min;
  }
}

/* ------------------------------------------------------------------------------------------------
 * NMS.  Sort + sweep + return order follow pet/lib/ops/csrc/NMS/ml_nms.cu:92-94,127-140,143-145;
 * IoU follows ml_nms.cu:11-26.  labels == NULL gives the class-agnostic `nms` the reference takes from
 * torchvision.ops.nms (pet/lib/ops/nms.py:2,10; torchvision 0.26, version un-pinned by the reference):
 * published algorithm = stable descending sort, suppress j>i when IoU(i,j) > thr, keep in sorted order.
 *
 * flavor selects how the fp32 union is rounded (the SOURCE is `Sa + Sb - inter` in all three):
 *   0  plain: every operation rounded separately  (C++ without contraction; == torchvision CPU)
 *   1  fmaf(bw, bh, Sa) - inter   (what torchvision 0.26's sm_100 nms kernel executes: SASS inspected)
 *   2  fmaf(aw, ah, Sb) - inter   (what nvcc 12.9 makes of ml_nms.cu:23-25 for sm_100a: SASS inspected)
 * where a = the higher-scored (row) box and b = the candidate (column) box.
 * Ties in score are broken by ascending input index (torchvision sorts stable; ml_nms.cu:92 does not
 * specify, so any order is admissible there).
 * ---------------------------------------------------------------------------------------------- */
static int orc_iou_gt(const float* a, const float* b, float thr, int flavor) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  float inter = w * h;
  float aw = a[2] - a[0], ah = a[3] - a[1], bw = b[2] - b[0], bh = b[3] - b[1];
  float uni;
  if (flavor == 1) { float Sa = aw * ah; uni = fmaf(bw, bh, Sa) - inter; }
  else if (flavor == 2) { float Sb = bw * bh; uni = fmaf(aw, ah, Sb) - inter; }
  else { float Sa = aw * ah, Sb = bw * bh; uni = Sa + Sb - inter; }
  return (inter / uni) > thr;
}

typedef struct { float s; int64_t i; } orc_si;
static int orc_cmp_desc(const void* pa, const void* pb) {
  const orc_si* a = (const orc_si*)pa; const orc_si* b = (const orc_si*)pb;
  /* NaN sorts as the greatest value (torch.sort semantics) */
  int an = a->s != a->s, bn = b->s != b->s;
  if (an != bn) return an ? -1 : 1;
  if (!an) { if (a->s > b->s) return -1; if (a->s < b->s) return 1; }
  return a->i < b->i ? -1 : (a->i > b->i ? 1 : 0);
}

int64_t orc_nms(const float* boxes, const float* scores, const int64_t* labels, int64_t N, float thr,
                int64_t topk, int flavor, int64_t* keep) {
  if (N == 0) return 0;
  orc_si* ord = (orc_si*)malloc(sizeof(orc_si) * N);
  uint8_t* removed = (uint8_t*)calloc(N, 1);
  for (int64_t i = 0; i < N; i++) { ord[i].s = scores[i]; ord[i].i = i; }
  qsort(ord, N, sizeof(orc_si), orc_cmp_desc);
  int64_t nk = 0;
  for (int64_t i = 0; i < N; i++) {
    if (removed[i]) continue;
    keep[nk++] = ord[i].i;
    if (nk == topk) break;                       /* ml_nms.cu:134 (topk == 0 never matches) */
    const float* a = boxes + 4 * ord[i].i;
    for (int64_t j = i + 1; j < N; j++) {
      if (removed[j]) continue;
      if (labels && labels[ord[i].i] != labels[ord[j].i]) continue;   /* ml_nms.cu:17 */
      if (orc_iou_gt(a, boxes + 4 * ord[j].i, thr, flavor)) removed[j] = 1;
    }
  }
  free(ord); free(removed);
  return nk;
}

/* ------------------------------------------------------------------------------------------------
 * Grid-point heat-map -> box decode.  pet/rcnn/modeling/grid_cascade_rcnn/inference.py:189-279
 * (GridPostProcessor.get_boxes); sub_xy = the (sub_x1, sub_y1) of calc_sub_regions
 * (pet/rcnn/modeling/grid_rcnn/loss.py:244-273).  logits (R,P,h,w); boxes (R,4); out (R,4).
 * The reference takes the arg-max over sigmoid(logits) (first index on ties, torch CPU max(dim)),
 * which is what is done here; `scores_out`/`pos_out` (R,P) are optional debug outputs.
 * The trailing clamp_ at :275-276 acts on a copy and is a no-op, so boxes are returned un-clamped.
 * ---------------------------------------------------------------------------------------------- */
void orc_grid_decode(const float* logits, const float* boxes, int R, int P, int h, int w, const int* sub_xy,
                     float mapping_ratio, float* out, float* scores_out, int* pos_out) {
  int gs = (int)(sqrt((double)P) + 0.5);
  float* sc = (float*)malloc(sizeof(float) * P);
  float* ax = (float*)malloc(sizeof(float) * P);
  float* ay = (float*)malloc(sizeof(float) * P);
  for (int r = 0; r < R; r++) {
    const float* b = boxes + 4 * r;
    float width = b[2] - b[0], height = b[3] - b[1];
    float x1 = b[0] - mapping_ratio * (width / 2), y1 = b[1] - mapping_ratio * (height / 2);
    for (int p = 0; p < P; p++) {
      const float* m = logits + ((size_t)r * P + p) * h * w;
      float best = -1.f; int bi = 0;
      for (int i = 0; i < h * w; i++) {
        float s = 1.0f / (1.0f + expf(-m[i]));
        if (s > best) { best = s; bi = i; }
      }
      int xs = bi % w + sub_xy[2 * p], ys = bi / w + sub_xy[2 * p + 1];
      sc[p] = best;
      ax[p] = ((float)xs + 0.5f) / (float)(2 * w) * (1.f + mapping_ratio) * width + x1;
      ay[p] = ((float)ys + 0.5f) / (float)(2 * h) * (1.f + mapping_ratio) * height + y1;
      if (scores_out) scores_out[r * P + p] = best;
      if (pos_out) pos_out[r * P + p] = bi;
    }
    float nx1 = 0, dx1 = 0, ny1 = 0, dy1 = 0, nx2 = 0, dx2 = 0, ny2 = 0, dy2 = 0;
    for (int i = 0; i < gs; i++) {
      int ix1 = i, iy1 = i * gs, ix2 = P - gs + i, iy2 = (i + 1) * gs - 1;   /* inference.py:251-258 */
      nx1 += ax[ix1] * sc[ix1]; dx1 += sc[ix1];
      ny1 += ay[iy1] * sc[iy1]; dy1 += sc[iy1];
      nx2 += ax[ix2] * sc[ix2]; dx2 += sc[ix2];
      ny2 += ay[iy2] * sc[iy2]; dy2 += sc[iy2];
    }
    out[4 * r + 0] = nx1 / dx1; out[4 * r + 1] = ny1 / dy1;
    out[4 * r + 2] = nx2 / dx2; out[4 * r + 3] = ny2 / dy2;
  }
  free(sc); free(ax); free(ay);
}
