/*
 * simsense_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement (plain C + OpenMP over rows/columns) of the algorithm of the
 * reference simsense DepthSensorEngine (reference: 3rd_party/simsense/src/).  It is the
 * checker for the CUDA path in sapien_b200/csrc: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference-cpu leg may load it.  The product never does.
 *
 * Parity pin: the reference has no golden vectors for this path.  The oracle is
 * pinned against the reference's own CUDA kernels, recompiled unmodified for sm_100a
 * (oracle/Makefile -> oracle/_ref/libsimsense_ref.so) and run on a B200; the per-stage outputs
 * of that run are committed under tests/golden/ (tests/golden/make_golden.py) and
 * tests/test_oracle_golden.py checks every function below against them.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/3rd_party/simsense/).  Volumes are uint16 [rows][cols][D], D innermost.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define IDX3(y, x, d) ((((size_t)(y)) * cols + (x)) * D + (d))

/* src/core.cu:45-62 float2uint8: red channel of RGBA f32, trunc(v*255), clamp to [0,255]. */
void orc_float2uint8(const float *rgba, uint8_t *dst, int rows, int cols) {
  const long n = (long)rows * cols;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; ++p) {
    float f = rgba[4 * p] * 255.0f;
    int t;
    if (!(f == f)) t = 0;                 /* CUDA float->int of NaN is 0 */
    else if (f >= 2147483648.0f) t = INT_MAX; /* CUDA saturates */
    else if (f <= -2147483648.0f) t = INT_MIN;
    else t = (int)f;
    dst[p] = (uint8_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
  }
}

/* src/camera.cu:77-120 remap.  The predicate at :87/:90 is always true, so the map coordinate is
 * always snapped with round-half-away, clamped, and the bilinear blend degenerates to the single
 * texel (x1,y1): nearest-neighbour gather. */
void orc_remap(const float *mapx, const float *mapy, const uint8_t *src, uint8_t *dst, int rows,
               int cols) {
  const long n = (long)rows * cols;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; ++p) {
    float sx = roundf(mapx[p]);
    float sy = roundf(mapy[p]);
    if (sx < 0) sx = 0;
    if (sx > cols - 1) sx = (float)(cols - 1);
    if (sy < 0) sy = 0;
    if (sy > rows - 1) sy = (float)(rows - 1);
    dst[p] = src[(long)sy * cols + (long)sx];
  }
}

/* src/camera.cu:122-139 copySubArea. */
void orc_crop_u8(const uint8_t *src, uint8_t *dst, int srcW, int bw, int bh, int sx, int sy) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < bh; ++y)
    memcpy(dst + (size_t)y * bw, src + (size_t)(sy + y) * srcW + sx, (size_t)bw);
}

/* src/camera.cu:141-158 pasteSubArea. */
void orc_paste_f32(const float *src, float *dst, int dstW, int bw, int bh, int sx, int sy) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < bh; ++y)
    memcpy(dst + (size_t)(sy + y) * dstW + sx, src + (size_t)y * bw, sizeof(float) * (size_t)bw);
}

/* src/csct.cu:21-87 CSCT: centre-symmetric census, zero padding outside the image. */
static inline int px0(const uint8_t *im, int rows, int cols, int y, int x) {
  return (x < 0 || x >= cols || y < 0 || y >= rows) ? 0 : im[(size_t)y * cols + x];
}
void orc_census(const uint8_t *im, uint32_t *census, int rows, int cols, int cw, int ch) {
  const int left = (cw - 1) / 2, top = (ch - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      uint32_t r = 0;
      for (int i = 0; i < top + 1; ++i) {
        const int jmax = (i == top) ? cw / 2 : cw;
        for (int j = 0; j < jmax; ++j) {
          const int a = px0(im, rows, cols, y - top + i, x - left + j);
          const int b = px0(im, rows, cols, y + top - i, x + left - j);
          const int sh = i * cw + j;
          /* a 32-bit shift by >=32 yields 0 on the GPU (shl clamps); cannot happen for cw*ch<=65 */
          if (sh < 32) r |= ((uint32_t)(a >= b)) << sh;
        }
      }
      census[(size_t)y * cols + x] = r;
    }
}

/* src/cost.cu:21-47 hammingCost: C0(y,x,d)=popc(L(y,x)^R(y,max(x-d,0))). */
void orc_hamming(const uint32_t *cl, const uint32_t *cr, uint16_t *cost, int rows, int cols,
                 int D) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const uint32_t base = cl[(size_t)y * cols + x];
      for (int d = 0; d < D; ++d) {
        const int xr = x - d < 0 ? 0 : x - d;
        cost[IDX3(y, x, d)] = (uint16_t)__builtin_popcount(base ^ cr[(size_t)y * cols + xr]);
      }
    }
}

/* src/filter.cu:49-70 boxFilterHorizontal: running sum with replicate borders, uint16 arithmetic. */
void orc_box_h(const uint16_t *in, uint16_t *out, int rows, int cols, int D, int size) {
  const int half = size / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int d = 0; d < D; ++d) {
      uint16_t acc = 0;
      for (int i = 0; i <= half; ++i) {
        const int scale = (i == 0) ? half + 1 : 1;
        acc = (uint16_t)(acc + in[IDX3(y, i, d)] * scale);
      }
      out[IDX3(y, 0, d)] = acc;
      for (int x = 1; x < cols; ++x) {
        const int xa = x + half < cols - 1 ? x + half : cols - 1;
        const int xs = x - (half + 1) > 0 ? x - (half + 1) : 0;
        acc = (uint16_t)(acc + in[IDX3(y, xa, d)] - in[IDX3(y, xs, d)]);
        out[IDX3(y, x, d)] = acc;
      }
    }
}

/* src/filter.cu:75-96 boxFilterVertical. */
void orc_box_v(const uint16_t *in, uint16_t *out, int rows, int cols, int D, int size) {
  const int half = size / 2;
#pragma omp parallel for schedule(static)
  for (int x = 0; x < cols; ++x)
    for (int d = 0; d < D; ++d) {
      uint16_t acc = 0;
      for (int i = 0; i <= half; ++i) {
        const int scale = (i == 0) ? half + 1 : 1;
        acc = (uint16_t)(acc + in[IDX3(i, x, d)] * scale);
      }
      out[IDX3(0, x, d)] = acc;
      for (int y = 1; y < rows; ++y) {
        const int ya = y + half < rows - 1 ? y + half : rows - 1;
        const int ys = y - (half + 1) > 0 ? y - (half + 1) : 0;
        acc = (uint16_t)(acc + in[IDX3(ya, x, d)] - in[IDX3(ys, x, d)]);
        out[IDX3(y, x, d)] = acc;
      }
    }
}

/* One SGM step, src/aggr.cu:39-76 (identical in all four kernels): int32 math, uint16 store. */
static inline void sgm_step(const uint16_t *c, const uint16_t *prev, uint16_t *out, int D, int P1,
                            int P2) {
  int m = INT_MAX;
  for (int d = 0; d < D; ++d)
    if (prev[d] < m) m = prev[d];
  for (int d = 0; d < D; ++d) {
    int best = prev[d];
    if (d != 0 && prev[d - 1] + P1 < best) best = prev[d - 1] + P1;
    if (d != D - 1 && prev[d + 1] + P1 < best) best = prev[d + 1] + P1;
    if (m + P2 < best) best = m + P2;
    out[d] = (uint16_t)(c[d] + best - m);
  }
}

/* src/aggr.cu:29-77 (dir 0, left->right), :79-127 (dir 1, right->left), :129-177 (dir 2,
 * top->bottom), :179-230 (dir 3, bottom->top; the /4 blend is orc_sum4 below). */
void orc_aggr(const uint16_t *cost, uint16_t *L, int dir, int P1, int P2, int rows, int cols,
              int D) {
  if (dir == 0 || dir == 1) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y) {
      const int x0 = dir == 0 ? 0 : cols - 1, dx = dir == 0 ? 1 : -1;
      memcpy(L + IDX3(y, x0, 0), cost + IDX3(y, x0, 0), sizeof(uint16_t) * (size_t)D);
      for (int k = 1, x = x0 + dx; k < cols; ++k, x += dx)
        sgm_step(cost + IDX3(y, x, 0), L + IDX3(y, x - dx, 0), L + IDX3(y, x, 0), D, P1, P2);
    }
  } else {
#pragma omp parallel for schedule(static)
    for (int x = 0; x < cols; ++x) {
      const int y0 = dir == 2 ? 0 : rows - 1, dy = dir == 2 ? 1 : -1;
      memcpy(L + IDX3(y0, x, 0), cost + IDX3(y0, x, 0), sizeof(uint16_t) * (size_t)D);
      for (int k = 1, y = y0 + dy; k < rows; ++k, y += dy)
        sgm_step(cost + IDX3(y, x, 0), L + IDX3(y - dy, x, 0), L + IDX3(y, x, 0), D, P1, P2);
    }
  }
}

/* src/aggr.cu:192,222: LAll=(L3+L0+L1+L2)/4, int floor, uint16 store. */
void orc_sum4(const uint16_t *L0, const uint16_t *L1, const uint16_t *L2, const uint16_t *L3,
              uint16_t *LAll, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    LAll[i] = (uint16_t)(((int)L3[i] + (int)L0[i] + (int)L1[i] + (int)L2[i]) / 4);
}

/* src/wta.cu:170-214 winnerTakesAll (reducers :30-65, :139-162 resolve ties to the lowest d;
 * sub-pixel :164-168 is evaluated in double then narrowed to float). */
void orc_wta(const uint16_t *LAll, float *dispL, uint16_t *dispR, int rows, int cols, int D,
             int uniq) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const uint16_t *v = LAll + IDX3(y, x, 0);
      int mv = INT_MAX, mi = 0;
      for (int d = 0; d < D; ++d)
        if (v[d] < mv) { mv = v[d]; mi = d; }
      int rv = INT_MAX, ri = 0;
      for (int d = 0; d < D && x + d < cols; ++d) {
        const int c = LAll[IDX3(y, x + d, d)];
        if (c < rv) { rv = c; ri = d; }
      }
      dispR[(size_t)y * cols + x] = (uint16_t)ri;
      int unique = 1;
      for (int d = 0; d < D; ++d)
        if (!((int)v[d] * (100 - uniq) >= mv * 100 || abs(mi - d) <= 1)) { unique = 0; break; }
      float out = (float)mi;
      if (!unique) out = -1.0f;
      else if (mi != 0 && mi != D - 1) {
        const int y0 = v[mi - 1], y1 = mv, y2 = v[mi + 1];
        const float a = (float)((1.0 * (y2 - y0)) / (2.0 * (y0 - 2 * y1 + y2)));
        out = (float)mi - a;
      }
      dispL[(size_t)y * cols + x] = out;
    }
}

/* src/lrcheck.cu:21-32 lrConsistencyCheck (in place). */
void orc_lrcheck(float *dispL, const uint16_t *dispR, int rows, int cols, int lrMaxDiff) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const size_t p = (size_t)y * cols + x;
      const int ld = (int)roundf(dispL[p]);
      if (ld < 0 || x - ld < 0 || abs(ld - (int)dispR[p - ld]) > lrMaxDiff) dispL[p] = -1.0f;
    }
}

/* src/filter.cu:98-117 medianFilter (+ getMedian :21-44): interior pixels get the (k*k/2)-th
 * order statistic, border pixels are copied; -1 takes part as an ordinary value. */
static int cmp_f32(const void *a, const void *b) {
  const float fa = *(const float *)a, fb = *(const float *)b;
  return (fa > fb) - (fa < fb);
}
void orc_median(const float *in, float *out, int rows, int cols, int size) {
  const int h = size / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y) {
    float win[49];
    for (int x = 0; x < cols; ++x) {
      const size_t p = (size_t)y * cols + x;
      if (x >= h && y >= h && x < cols - h && y < rows - h) {
        for (int j = 0; j < size; ++j)
          for (int i = 0; i < size; ++i)
            win[j * size + i] = in[(size_t)(y - h + j) * cols + x - h + i];
        qsort(win, (size_t)size * size, sizeof(float), cmp_f32);
        out[p] = win[size * size / 2];
      } else {
        out[p] = in[p];
      }
    }
  }
}

/* src/camera.cu:160-168 disp2Depth: (f*b)/disp, product first; disp<=0 -> 0. */
void orc_disp2depth(const float *disp, float *depth, long n, float focal, float baseline) {
  const float fb = focal * baseline;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; ++p) depth[p] = (disp[p] <= 0) ? 0.0f : fb / disp[p];
}

static inline int f2i_sat(float f) { /* CUDA cvt.rzi.s32.f32: NaN->0, saturating */
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return INT_MAX;
  if (f <= -2147483648.0f) return INT_MIN;
  return (int)f;
}

/* src/camera.cu:170-196 initRgbDepth + depthRegistration (atomicMinFloat :42-47), :200-228
 * depthDilation with snapshot semantics, :230-240 correctDepthRange.
 * nvcc contracts a*z+b to an FMA, hence fmaf. */
void orc_register(const float *depth, const float *a1, const float *a2, const float *a3, float b1,
                  float b2, float b3, int rows, int cols, float *rgbDepth, float *scratch,
                  int rgbRows, int rgbCols, int dilation, float minDepth, float maxDepth) {
  const long rn = (long)rgbRows * rgbCols;
  const long n = (long)rows * cols;
  float *canvas = dilation ? scratch : rgbDepth;
  for (long p = 0; p < rn; ++p) canvas[p] = maxDepth;
  for (long p = 0; p < n; ++p) { /* min is order independent; serial keeps it race-free */
    const float z = depth[p];
    const float zr = fmaf(a3[p], z, b3);
    const int u = f2i_sat(roundf(fmaf(a1[p], z, b1) / zr));
    const int v = f2i_sat(roundf(fmaf(a2[p], z, b2) / zr));
    if (zr > 0 && u >= 0 && u < rgbCols && v >= 0 && v < rgbRows) {
      float *q = canvas + (long)v * rgbCols + u;
      if (zr < *q) *q = zr;
    }
  }
  if (dilation) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rgbRows; ++y)
      for (int x = 0; x < rgbCols; ++x) {
        float m = canvas[(long)y * rgbCols + x];
        const int xr = x + 1 < rgbCols, yb = y + 1 < rgbRows;
        float t;
        if (xr && (t = canvas[(long)y * rgbCols + x + 1]) < maxDepth && t < m) m = t;
        if (yb && (t = canvas[(long)(y + 1) * rgbCols + x]) < maxDepth && t < m) m = t;
        if (xr && yb && (t = canvas[(long)(y + 1) * rgbCols + x + 1]) < maxDepth && t < m) m = t;
        rgbDepth[(long)y * rgbCols + x] = m;
      }
  }
#pragma omp parallel for schedule(static)
  for (long p = 0; p < rn; ++p)
    if (rgbDepth[p] < minDepth || rgbDepth[p] >= maxDepth) rgbDepth[p] = 0.0f;
}

/* src/camera.cu:230-240 correctDepthRange alone (non-registration constructor path). */
void orc_range(float *depth, long n, float minDepth, float maxDepth) {
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; ++p)
    if (depth[p] < minDepth || depth[p] >= maxDepth) depth[p] = 0.0f;
}

/* src/camera.cu:242-261 depth2PointCloud / :263-286 depth2RgbPointCloud (rgba may be NULL). */
void orc_pointcloud(const float *depth, const float *rgba, float *pc, int rows, int cols, float fx,
                    float fy, float s, float cx, float cy) {
  const int stride = rgba ? 6 : 3;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      const long p = (long)v * cols + u;
      const float z = depth[p];
      const float t1 = ((float)u - cx) / fx;
      const float t2 = (s * (cy - (float)v)) / (fx * fy);
      const float x = z * (t1 + t2);
      const float y = (z * ((float)v - cy)) / fy;
      pc[stride * p] = x;
      pc[stride * p + 1] = y;
      pc[stride * p + 2] = z;
      if (rgba) {
        pc[stride * p + 3] = rgba[4 * p];
        pc[stride * p + 4] = rgba[4 * p + 1];
        pc[stride * p + 5] = rgba[4 * p + 2];
      }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Whole pipeline, sequenced as src/core.cu:543-787 computeDepth.  Optional stage outputs may be
 * NULL.  Outside-ROI disparity is defined as 0 (the reference never clears d_bboxDisp,
 * core.cu:283).                                                          */
typedef struct {
  int rows, cols, rgb_rows, rgb_cols;
  float focal, baseline, min_depth, max_depth;
  int rectified, census_w, census_h, max_disp, bf_w, bf_h, p1, p2, uniq, lr_max_diff, mf_size;
  float b1, b2, b3;
  int dilation, registration;
  int bbox, bbox_x, bbox_y, bbox_w, bbox_h;
} orc_cfg;

typedef struct { /* all optional; sized for the matched (ROI) image unless noted */
  uint8_t *im0, *im1;          /* images fed to census (after remap / crop)          */
  uint32_t *census0, *census1; /*                                                    */
  uint16_t *rawcost, *cost, *L0, *L1, *L2, *L3, *LAll;
  float *disp_wta, *disp_lr, *disp_med; /* left disparity after WTA / LR / median    */
  uint16_t *disp_right;
  float *disp_full;            /* full IR size, after paste                          */
  float *depth;                /* full IR size                                        */
} orc_stages;

#define KEEP(dst, src, bytes) do { if (dst) memcpy(dst, src, bytes); } while (0)

int orc_pipeline(const orc_cfg *c, const uint8_t *left, const uint8_t *right, const float *mapLx,
                 const float *mapLy, const float *mapRx, const float *mapRy, const float *a1,
                 const float *a2, const float *a3, float *out_depth, orc_stages *st) {
  const int fr = c->rows, fc = c->cols;
  const size_t fsz = (size_t)fr * fc;
  uint8_t *s0 = (uint8_t *)malloc(fsz), *s1 = (uint8_t *)malloc(fsz);
  const uint8_t *i0 = left, *i1 = right;
  if (!c->rectified) {
    orc_remap(mapLx, mapLy, left, s0, fr, fc);
    orc_remap(mapRx, mapRy, right, s1, fr, fc);
    i0 = s0; i1 = s1;
  }
  int rows = fr, cols = fc;
  uint8_t *b0 = NULL, *b1 = NULL;
  if (c->bbox) {
    rows = c->bbox_h; cols = c->bbox_w;
    b0 = (uint8_t *)malloc((size_t)rows * cols); b1 = (uint8_t *)malloc((size_t)rows * cols);
    orc_crop_u8(i0, b0, fc, cols, rows, c->bbox_x, c->bbox_y);
    orc_crop_u8(i1, b1, fc, cols, rows, c->bbox_x, c->bbox_y);
    i0 = b0; i1 = b1;
  }
  const size_t sz = (size_t)rows * cols;
  const int D = c->max_disp;
  const size_t vsz = sz * D;
  if (st) { KEEP(st->im0, i0, sz); KEEP(st->im1, i1, sz); }
  uint32_t *c0 = (uint32_t *)malloc(4 * sz), *c1 = (uint32_t *)malloc(4 * sz);
  orc_census(i0, c0, rows, cols, c->census_w, c->census_h);
  orc_census(i1, c1, rows, cols, c->census_w, c->census_h);
  if (st) { KEEP(st->census0, c0, 4 * sz); KEEP(st->census1, c1, 4 * sz); }
  uint16_t *cost = (uint16_t *)malloc(2 * vsz);
  uint16_t *A = (uint16_t *)malloc(2 * vsz), *B = (uint16_t *)malloc(2 * vsz);
  if (c->bf_w * c->bf_h == 1) {
    orc_hamming(c0, c1, cost, rows, cols, D);
    if (st) KEEP(st->rawcost, cost, 2 * vsz);
  } else {
    orc_hamming(c0, c1, A, rows, cols, D);
    if (st) KEEP(st->rawcost, A, 2 * vsz);
    orc_box_h(A, B, rows, cols, D, c->bf_w);
    orc_box_v(B, cost, rows, cols, D, c->bf_h);
  }
  if (st) KEEP(st->cost, cost, 2 * vsz);
  const int P1 = c->p1 * c->bf_w * c->bf_h, P2 = c->p2 * c->bf_w * c->bf_h; /* core.cu:670-671 */
  uint16_t *acc = (uint16_t *)malloc(2 * vsz); /* holds L0, then the running blend inputs */
  uint16_t *L1 = (uint16_t *)malloc(2 * vsz), *L2 = (uint16_t *)malloc(2 * vsz);
  orc_aggr(cost, acc, 0, P1, P2, rows, cols, D);
  orc_aggr(cost, L1, 1, P1, P2, rows, cols, D);
  orc_aggr(cost, L2, 2, P1, P2, rows, cols, D);
  orc_aggr(cost, A, 3, P1, P2, rows, cols, D);
  if (st) { KEEP(st->L0, acc, 2 * vsz); KEEP(st->L1, L1, 2 * vsz); KEEP(st->L2, L2, 2 * vsz);
            KEEP(st->L3, A, 2 * vsz); }
  orc_sum4(acc, L1, L2, A, B, vsz);
  if (st) KEEP(st->LAll, B, 2 * vsz);
  float *dl = (float *)malloc(4 * sz), *df = (float *)malloc(4 * sz);
  uint16_t *dr = (uint16_t *)malloc(2 * sz);
  orc_wta(B, dl, dr, rows, cols, D, c->uniq);
  if (st) { KEEP(st->disp_wta, dl, 4 * sz); KEEP(st->disp_right, dr, 2 * sz); }
  if (c->lr_max_diff != 255) orc_lrcheck(dl, dr, rows, cols, c->lr_max_diff);
  if (st) KEEP(st->disp_lr, dl, 4 * sz);
  float *disp = dl;
  if (c->mf_size != 1) { orc_median(dl, df, rows, cols, c->mf_size); disp = df; }
  if (st) KEEP(st->disp_med, disp, 4 * sz);
  float *full = disp, *canvas = NULL;
  if (c->bbox) {
    canvas = (float *)calloc(fsz, sizeof(float));
    orc_paste_f32(disp, canvas, fc, cols, rows, c->bbox_x, c->bbox_y);
    full = canvas;
  }
  if (st) KEEP(st->disp_full, full, 4 * fsz);
  float *depth = (float *)malloc(4 * fsz);
  orc_disp2depth(full, depth, (long)fsz, c->focal, c->baseline);
  if (st) KEEP(st->depth, depth, 4 * fsz);
  if (c->registration) {
    const size_t rsz = (size_t)c->rgb_rows * c->rgb_cols;
    float *scratch = (float *)malloc(4 * rsz);
    orc_register(depth, a1, a2, a3, c->b1, c->b2, c->b3, fr, fc, out_depth, scratch, c->rgb_rows,
                 c->rgb_cols, c->dilation, c->min_depth, c->max_depth);
    free(scratch);
  } else {
    memcpy(out_depth, depth, 4 * fsz);
    orc_range(out_depth, (long)fsz, c->min_depth, c->max_depth);
  }
  free(s0); free(s1); free(b0); free(b1); free(c0); free(c1); free(cost); free(A); free(B);
  free(acc); free(L1); free(L2); free(dl); free(df); free(dr); free(canvas); free(depth);
  return 0;
}
