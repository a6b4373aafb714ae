// front.cu -- image front-end for sm_100a: input conversion, optional IR noise, rectification
// gather, ROI crop and centre-symmetric census for BOTH images in one kernel.
//
// Replaces float2uint8 (3rd_party/simsense/src/core.cu:45-62), simInfraredNoise
// (src/camera.cu:58-75), remap (src/camera.cu:77-120, which is a nearest-neighbour gather because
// its snap predicate is always true -- SURVEY.md App. A-2), copySubArea (src/camera.cu:122-139)
// and CSCT (src/csct.cu:21-87).  The reference runs these as up to 9 launches with a device-wide
// sync between each and 32-thread blocks for the per-pixel ones; here a block stages the
// (32+cw-1) x (32+ch-1) census window of each image in shared memory straight from the source
// (through the map and the ROI offset), so no intermediate image round-trips through HBM.
#include "common.cuh"
#include "kernels.h"

namespace ssb {

constexpr int FT = 32; // output tile edge
constexpr int FTY = 8; // threads along y (each thread produces FT/FTY rows)

// ---- counter-based RNG for the IR noise model (stateless replacement for the reference's
// 48 B/pixel XORWOW states, camera.cu:49-56).  Philox-4x32-10. ----------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
struct Philox {
  uint32_t ctr[4], key[2];
  __device__ Philox(uint64_t seed, uint64_t pixel, uint64_t frame) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = (uint32_t)frame; ctr[2] = (uint32_t)pixel; ctr[3] = (uint32_t)(pixel >> 32) ^ (uint32_t)(frame >> 32);
  }
  // one block = four 32-bit outputs; successive calls advance the counter
  __device__ void block(uint32_t (&out)[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    ctr[0]++;
  }
};
__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; } // (0,1)

// IR speckle + thermal noise of one source texel (camera.cu:21-40,66-74):
//   round(I * Gamma(shape, scale) + mu + sigma * N(0,1)), clamped to [0,255],
// Gamma by Marsaglia-Tsang exactly as the reference (boost by u^(1/shape) for shape < 1).  One Philox block serves
// the common case: its first two outputs give BOTH Box-Muller normals (the cosine one proposes the Gamma variate, the
// sine one is the thermal noise -- the two are independent), the third the acceptance uniform; only a rejected proposal
// (~2 % at the stock shape) or shape < 1 draws further blocks.  The acceptance test is the reference's, preceded by
// Marsaglia-Tsang's squeeze (u < 1 - 0.0331 z^4 implies acceptance), which skips both logarithms most of the time
// without changing a single decision.
__device__ __noinline__ int ir_noise(const FrontParams &p, int v, size_t sp, int which) {
  Philox g(p.seed + (uint64_t)which, (uint64_t)sp, p.frame);
  uint32_t r[4];
  g.block(r);
  float rad = sqrtf(-2.0f * logf(u01(r[0])));
  float sn, cs;
  sincospif(2.0f * u01(r[1]), &sn, &cs);
  const float thermal = rad * sn;
  float z = rad * cs, u = u01(r[2]);
  float boost = 1.0f, alpha = p.speckle_shape;
  if (alpha < 1.0f) {
    boost = powf(u01(r[3]), 1.0f / alpha);
    alpha += 1.0f;
  }
  const float d = alpha - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
  float gam = d; // (after 64 rejections in a row -- probability ~1e-100 -- the mode)
  for (int it = 0; it < 64; ++it) {
    const float t = 1.0f + c * z;
    const float w = t * t * t;
    const float z2 = z * z;
    if (t > 0.0f && (u < 1.0f - 0.0331f * z2 * z2 || logf(u) < 0.5f * z2 + d - d * w + d * logf(w))) { gam = d * w; break; }
    g.block(r);
    rad = sqrtf(-2.0f * logf(u01(r[0])));
    z = rad * cospif(2.0f * u01(r[1]));
    u = u01(r[2]);
  }
  const float res = roundf((float)v * (gam * p.speckle_scale * boost) + p.gaussian_mu + p.gaussian_sigma * thermal);
  return min(max((int)res, 0), 255);
}

// noise stream id of a texel = its PACKED pixel index inside the environment, whatever the source pitch
template <bool RGBA> __device__ __noinline__ uint32_t noise_pixel(const FrontParams &p, uint32_t off) {
  const uint32_t yy = off / p.src_row, xx = (off - yy * p.src_row) / (RGBA ? 4u : 1u);
  return yy * (uint32_t)p.fcols + xx;
}

struct Src {
  const uint8_t *u8;
  const float *rgba;
  const float *mapx, *mapy;
};

// Rectification map of pixel (u, v) from matrices (ss_create_calibrated): what cv2.initUndistortRectifyMap(K, None,
// R, P, size, CV_32F) tabulates -- [X,Y,W] = (P[:3,:3] R)^-1 [u,v,1]^T, map = K (X/W, Y/W) -- in float64, rounded to
// float32 like the CV_32F planes.
__device__ __forceinline__ void cal_map(const FrontParams &p, int img, int u, int v, float &mx, float &my) {
  const double *m = img ? p.rinvR : p.rinvL;
  const double du = (double)u, dv = (double)v;
  const double X = __dadd_rn(__dadd_rn(__dmul_rn(du, m[0]), __dmul_rn(dv, m[1])), m[2]);
  const double Y = __dadd_rn(__dadd_rn(__dmul_rn(du, m[3]), __dmul_rn(dv, m[4])), m[5]);
  const double W = __dadd_rn(__dadd_rn(__dmul_rn(du, m[6]), __dmul_rn(dv, m[7])), m[8]);
  const double iw = 1.0 / W;
  mx = (float)__fma_rn(p.ir_fx, X * iw, p.ir_cx);
  my = (float)__fma_rn(p.ir_fy, Y * iw, p.ir_cy);
}

template <bool RGBA>
__device__ __forceinline__ int fetch(const FrontParams &p, const Src &s, int n, int y, int x,
                                     int which) {
  if (x < 0 || x >= p.cols || y < 0 || y >= p.rows) return 0; // zero padding (csct.cu:40-42)
  int fx = x + p.bx, fy = y + p.by;
  if (s.mapx || p.cal_maps) {
    float mx, my;
    if (p.cal_maps) {
      cal_map(p, which, fx, fy, mx, my);
    } else {
      const size_t mp = (size_t)fy * p.fcols + fx;
      mx = __ldg(s.mapx + mp); my = __ldg(s.mapy + mp);
    }
    float sx = roundf(mx), sy = roundf(my);
    sx = fminf(fmaxf(sx, 0.0f), (float)(p.fcols - 1));
    sy = fminf(fmaxf(sy, 0.0f), (float)(p.frows - 1));
    fx = (int)sx; fy = (int)sy;
  }
  const size_t so = (size_t)n * p.src_env + (size_t)fy * p.src_row + (RGBA ? 4 : 1) * (size_t)fx; // pitched source
  int v;
  if (RGBA) {
    const int t = (int)(__ldg(s.rgba + so) * 255); // truncation, core.cu:51
    v = min(max(t, 0), 255);
  } else {
    v = __ldg(s.u8 + so);
  }
  if (p.speckle_shape > 0.0f) v = ir_noise(p, v, ((size_t)n * p.frows + fy) * p.fcols + fx, which);
  return v;
}

template <bool RGBA>
__global__ void __launch_bounds__(FT *FTY) front_kernel(const FrontParams p) {
  extern __shared__ uint8_t win[];
  const int wc = FT + p.cw - 1, wr = FT + p.ch - 1;
  uint8_t *w0 = win, *w1 = win + wc * wr;
  const int left = (p.cw - 1) / 2, top = (p.ch - 1) / 2;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * FT, y0 = blockIdx.y * FT;
  const int tid = threadIdx.y * FT + threadIdx.x;
  if (p.canvas) { // grid-stride fill of the registration canvas (float4 stores; tail by scalars)
    const size_t nthr = (size_t)gridDim.x * gridDim.y * gridDim.z * (FT * FTY);
    const size_t gtid = (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (FT * FTY) + tid;
    const size_t n4 = p.canvas_n / 4;
    const float4 f4 = make_float4(p.canvas_fill, p.canvas_fill, p.canvas_fill, p.canvas_fill);
    for (size_t i = gtid; i < n4; i += nthr) reinterpret_cast<float4 *>(p.canvas)[i] = f4;
    for (size_t i = n4 * 4 + gtid; i < p.canvas_n; i += nthr) p.canvas[i] = p.canvas_fill;
  }
  const Src sl{p.left_u8, p.left_rgba, p.mapLx, p.mapLy};
  const Src sr{p.right_u8, p.right_rgba, p.mapRx, p.mapRy};
  for (int i = tid; i < wc * wr; i += FT * FTY) {
    const int wy = i / wc, wx = i - wy * wc;
    const int y = y0 + wy - top, x = x0 + wx - left;
    w0[i] = (uint8_t)fetch<RGBA>(p, sl, n, y, x, 0);
    w1[i] = (uint8_t)fetch<RGBA>(p, sr, n, y, x, 1);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x;
  if (x >= p.cols) return;
#pragma unroll
  for (int ry = 0; ry < FT / FTY; ++ry) {
    const int ty = threadIdx.y + ry * FTY;
    const int y = y0 + ty;
    if (y >= p.rows) break;
    uint32_t r0 = 0, r1 = 0;
    for (int i = 0; i <= top; ++i) {
      const int jmax = (i == top) ? p.cw / 2 : p.cw;
      const uint8_t *a0 = w0 + (ty + i) * wc + threadIdx.x;
      const uint8_t *b0 = w0 + (ty + 2 * top - i) * wc + threadIdx.x + 2 * left;
      const uint8_t *a1 = w1 + (ty + i) * wc + threadIdx.x;
      const uint8_t *b1 = w1 + (ty + 2 * top - i) * wc + threadIdx.x + 2 * left;
      for (int j = 0; j < jmax; ++j) {
        const int sh = i * p.cw + j;
        const uint32_t bit = sh < 32 ? (1u << sh) : 0u; // shl by >=32 gives 0 on the GPU
        if (a0[j] >= b0[-j]) r0 |= bit;
        if (a1[j] >= b1[-j]) r1 |= bit;
      }
    }
    const size_t o = ((size_t)n * p.rows + y) * p.cols + x;
    p.census0[o] = r0;
    p.census1[o] = r1;
    p.im0[o] = w0[(ty + top) * wc + threadIdx.x + left];
    p.im1[o] = w1[(ty + top) * wc + threadIdx.x + left];
  }
}

// ---- 7x7 census fast path ------------------------------------------------------------------------
// Tile of 64 x 32 pixels per 128-thread block; a thread produces the codes of 4 adjacent pixels in 4
// rows for both images.  The window is staged as BYTES (one fetch through map / ROI / conversion per
// window texel, fully unrolled so the dependent map -> texel loads of all 22 rounds are in flight
// together), and the 24 comparisons of csct.cu:61-86 run 4 pixels at a time on packed bytes:
// ge(a,b) per byte lands in bit 7, and "acc = acc >> 1 | ge & 0x80808080" collects 8 consecutive
// code bits per byte lane, i.e. one byte of the code of each of the 4 pixels.
constexpr int F7_TW = 64, F7_TH = 32;
constexpr int F7_WB = F7_TW + 8; // staged bytes per window row: columns X0-4 .. X0+67
constexpr int F7_WH = F7_TH + 6; // staged rows: Y0-3 .. Y0+34
constexpr int F7_NT = 128;

__device__ __forceinline__ uint32_t bytes_ge_bit7(uint32_t a, uint32_t b) {
  // per byte: bit 7 = (a >= b); other bits unspecified.  (a|H)-(b&~H) cannot borrow across bytes.
  const uint32_t t = (a | 0x80808080u) - (b & 0x7f7f7f7fu);
  return (a & ~b) | (~(a ^ b) & t);
}
// 4 bytes starting at byte offset o (0..8) of the 12-byte group w0,w1,w2
template <int O> __device__ __forceinline__ uint32_t bytes_at(uint32_t w0, uint32_t w1, uint32_t w2) {
  if constexpr (O == 0) return w0;
  else if constexpr (O == 4) return w1;
  else if constexpr (O == 8) return w2;
  else if constexpr (O < 4) return __byte_perm(w0, w1, 0x3210 + 0x1111 * O);
  else return __byte_perm(w1, w2, 0x3210 + 0x1111 * (O - 4));
}

template <int I, int J> struct CensusBits {
  // comparisons (I, J..) of one output row in ascending code-bit order (bit = I*7 + J)
  static __device__ __forceinline__ void run(const uint32_t (&w)[7][3], uint32_t &acc, uint32_t (&grp)[3]) {
    constexpr int JMAX = (I == 3) ? 3 : 7;
    if constexpr (I <= 3 && J < JMAX) {
      // a = P(y-3+I, x-3+J), b = P(y+3-I, x+3-J); staged column 0 is x0-4, so pixel x0's byte offsets
      // are 1+J and 7-J
      const uint32_t a = bytes_at<1 + J>(w[I][0], w[I][1], w[I][2]);
      const uint32_t b = bytes_at<7 - J>(w[6 - I][0], w[6 - I][1], w[6 - I][2]);
      acc = ((acc >> 1) & 0x7f7f7f7fu) | (bytes_ge_bit7(a, b) & 0x80808080u); // one SHF + one LOP3
      constexpr int bit = I * 7 + J;
      if constexpr ((bit & 7) == 7) grp[bit >> 3] = acc;
      CensusBits<I, J + 1>::run(w, acc, grp);
    } else if constexpr (I < 3) {
      CensusBits<I + 1, 0>::run(w, acc, grp);
    }
  }
};

// One block = one tile of ONE image (blockIdx.z = 2 * env + image): twice the blocks and half the
// dependent load rounds per block of a both-images block (the kernel is bound by the latency of
// the map -> texel gather, not by bandwidth).
// PITCHED: the source rows / environments are not packed (FrontParams::src_row / src_env); the packed
// instantiation keeps its tighter address arithmetic (the kernel sits at its register cap).
// CALMAP: rectification maps evaluated from matrices (FrontParams::cal_maps) instead of read from planes.
template <bool RGBA, int MINB, bool PITCHED, bool CALMAP>
__global__ void __launch_bounds__(F7_NT, MINB) front7_kernel(const FrontParams p) {
  __shared__ __align__(16) uint32_t win[1][F7_WH][F7_WB / 4];
  const int tid = threadIdx.x;
  const int n = p.only_image < 0 ? blockIdx.z >> 1 : blockIdx.z;
  const int img = p.only_image < 0 ? blockIdx.z & 1 : p.only_image;
  const int X0 = blockIdx.x * F7_TW, Y0 = blockIdx.y * F7_TH;
  if (p.canvas) { // grid-stride fill of the registration canvas (initRgbDepth, camera.cu:170-177)
    const size_t nthr = (size_t)gridDim.x * gridDim.y * gridDim.z * F7_NT;
    const size_t gtid = (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * F7_NT + tid;
    const size_t n4 = p.canvas_n / 4;
    const float4 f4 = make_float4(p.canvas_fill, p.canvas_fill, p.canvas_fill, p.canvas_fill);
    for (size_t i = gtid; i < n4; i += nthr) reinterpret_cast<float4 *>(p.canvas)[i] = f4;
    for (size_t i = n4 * 4 + gtid; i < p.canvas_n; i += nthr) p.canvas[i] = p.canvas_fill;
  }
  const Src sl{p.left_u8, p.left_rgba, p.mapLx, p.mapLy};
  const Src sr{p.right_u8, p.right_rgba, p.mapRx, p.mapRy};
  uint8_t *b0 = reinterpret_cast<uint8_t *>(&win[0][0][0]);
  constexpr int NEL = F7_WB * F7_WH;
  constexpr int NRND = (NEL + F7_NT - 1) / F7_NT; // texels per thread and image
  constexpr int HALF = NRND; // one phase: every map load, then every texel load of the tile is in flight together
  constexpr int NPH = (NRND + HALF - 1) / HALF;
  // Branch-free staging in three straight-line phases (map loads -> texel loads -> convert/store) so
  // that all loads of a phase are in flight together; out-of-image texels are loaded from a clamped
  // address and replaced by the zero padding of csct.cu:40-42 afterwards.
  {
    const Src &s = img ? sr : sl;
    uint8_t *bdst = b0;
    const size_t envoff = (size_t)n * p.frows * p.fcols; // pixel id of the environment's first texel (noise streams)
    const size_t envsrc = PITCHED ? (size_t)n * p.src_env : (RGBA ? 4 : 1) * envoff;
    const float *srgba = RGBA ? s.rgba + envsrc : nullptr;
    const uint8_t *su8 = RGBA ? nullptr : s.u8 + envsrc;
#pragma unroll
    for (int h = 0; h < NPH; ++h) {
      int fx[HALF], fy[HALF];
      bool inside[HALF];
      float mx[HALF], my[HALF];
#pragma unroll
      for (int k = 0; k < HALF; ++k) {
        const int e = min((h * HALF + k) * F7_NT + tid, NEL - 1);
        const int wy = e / F7_WB, wx = e - wy * F7_WB;
        const int y = Y0 - 3 + wy, x = X0 - 4 + wx;
        inside[k] = x >= 0 && x < p.cols && y >= 0 && y < p.rows;
        fx[k] = min(max(x, 0), p.cols - 1) + p.bx;
        fy[k] = min(max(y, 0), p.rows - 1) + p.by;
        if (CALMAP) {
          cal_map(p, img, fx[k], fy[k], mx[k], my[k]);
        } else if (s.mapx) {
          const size_t mp = (size_t)fy[k] * p.fcols + fx[k];
          mx[k] = __ldg(s.mapx + mp);
          my[k] = __ldg(s.mapy + mp);
        }
      }
      float tf[HALF];
      uint8_t tb[HALF];
      uint32_t sp[HALF]; // packed: texel index inside environment n; PITCHED: element offset (one image is < 4 Gi elements)
#pragma unroll
      for (int k = 0; k < HALF; ++k) {
        if (CALMAP || s.mapx) { // camera.cu:83-119: always-snapped nearest neighbour
          const float sx = fminf(fmaxf(roundf(mx[k]), 0.0f), (float)(p.fcols - 1));
          const float sy = fminf(fmaxf(roundf(my[k]), 0.0f), (float)(p.frows - 1));
          fx[k] = (int)sx; fy[k] = (int)sy;
        }
        if (PITCHED) sp[k] = (uint32_t)fy[k] * p.src_row + (RGBA ? 4u : 1u) * (uint32_t)fx[k];
        else sp[k] = (uint32_t)fy[k] * (uint32_t)p.fcols + (uint32_t)fx[k];
        if (RGBA) tf[k] = __ldg(srgba + (PITCHED ? (size_t)sp[k] : 4 * (size_t)sp[k]));
        else tb[k] = __ldg(su8 + sp[k]);
      }
#pragma unroll
      for (int k = 0; k < HALF; ++k) {
        int v;
        if (RGBA) v = min(max((int)(tf[k] * 255), 0), 255); // truncation, core.cu:51
        else v = tb[k];
        if (p.speckle_shape > 0.0f) v = ir_noise(p, v, envoff + (PITCHED ? noise_pixel<RGBA>(p, sp[k]) : sp[k]), img);
        const int e = (h * HALF + k) * F7_NT + tid;
        if (e < NEL) bdst[e] = (uint8_t)(inside[k] ? v : 0);
      }
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4; // 16 x 8 threads, 4 pixels x 4 rows each
  const int x0 = X0 + 4 * tx;
  if (x0 >= p.cols) return;
  {
    uint32_t *census = img ? p.census1 : p.census0;
    uint8_t *im = img ? p.im1 : p.im0;
    uint32_t rows_w[10][3]; // window rows 4*ty .. 4*ty+9, words tx, tx+1, tx+2
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) rows_w[r][c] = win[0][4 * ty + r][tx + c];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = Y0 + 4 * ty + r;
      if (y >= p.rows) break;
      uint32_t w[7][3];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) w[i][c] = rows_w[r + i][c];
      }
      uint32_t acc = 0, grp[3] = {0, 0, 0};
      CensusBits<0, 0>::run(w, acc, grp);
      uint32_t code[4];
      code[0] = (__byte_perm(__byte_perm(grp[0], grp[1], 0x0040), grp[2], 0x0410)) & 0x00ffffffu;
      code[1] = (__byte_perm(__byte_perm(grp[0], grp[1], 0x0051), grp[2], 0x0510)) & 0x00ffffffu;
      code[2] = (__byte_perm(__byte_perm(grp[0], grp[1], 0x0062), grp[2], 0x0610)) & 0x00ffffffu;
      code[3] = (__byte_perm(__byte_perm(grp[0], grp[1], 0x0073), grp[2], 0x0710)) & 0x00ffffffu;
      const size_t o = ((size_t)n * p.rows + y) * p.cols + x0;
      const uint32_t centre = w[3][1]; // pixels x0..x0+3 of row y
      if (x0 + 3 < p.cols && (p.cols & 3) == 0) {
        *reinterpret_cast<uint4 *>(census + o) = make_uint4(code[0], code[1], code[2], code[3]);
        *reinterpret_cast<uint32_t *>(im + o) = centre;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (x0 + k < p.cols) { census[o + k] = code[k]; im[o + k] = (uint8_t)(centre >> (8 * k)); }
      }
    }
  }
}

cudaError_t launch_front(const FrontParams &p, cudaStream_t st) {
  if (p.N > 32767) return cudaErrorInvalidValue;
  if (p.cw == 7 && p.ch == 7) {
    const dim3 grid((p.cols + F7_TW - 1) / F7_TW, (p.rows + F7_TH - 1) / F7_TH, (p.only_image < 0 ? 2 : 1) * p.N);
    const bool packed = p.left_rgba ? (p.src_row == 4u * (uint32_t)p.fcols && p.src_env == 4 * (size_t)p.frows * p.fcols)
                                    : (p.src_row == (uint32_t)p.fcols && p.src_env == (size_t)p.frows * p.fcols);
    // (the pitched and the matrix-map variants are the general instantiation: <.., PITCHED = true, CALMAP>)
    if (p.left_rgba) {
      if (p.cal_maps) front7_kernel<true, 3, true, true><<<grid, F7_NT, 0, st>>>(p);
      // (three blocks per SM: 167 registers, no spills; four blocks = 128 registers spill a little: C4 +0.5 %, C1 +0.4 %;
      //  six / eight blocks spill more and are slower: C4 60.2 -> 58.9 / 57.4 k env-frames/s)
      else if (packed) front7_kernel<true, 3, false, false><<<grid, F7_NT, 0, st>>>(p);
      else front7_kernel<true, 3, true, false><<<grid, F7_NT, 0, st>>>(p);
    } else {
      if (p.cal_maps) front7_kernel<false, 3, true, true><<<grid, F7_NT, 0, st>>>(p);
      else if (packed) front7_kernel<false, 3, false, false><<<grid, F7_NT, 0, st>>>(p);
      else front7_kernel<false, 3, true, false><<<grid, F7_NT, 0, st>>>(p);
    }
    return cudaGetLastError();
  }
  if (p.only_image >= 0) return cudaErrorInvalidValue; // per-image launches: 7x7 census path only
  const dim3 grid((p.cols + FT - 1) / FT, (p.rows + FT - 1) / FT, p.N);
  const dim3 block(FT, FTY);
  const size_t smem = 2 * (size_t)(FT + p.cw - 1) * (FT + p.ch - 1);
  if (p.left_rgba) front_kernel<true><<<grid, block, smem, st>>>(p);
  else front_kernel<false><<<grid, block, smem, st>>>(p);
  return cudaGetLastError();
}

} // namespace ssb
