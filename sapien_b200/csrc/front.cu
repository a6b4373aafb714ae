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
  uint32_t ctr[4], key[2], out[4];
  int have;
  __device__ Philox(uint64_t seed, uint64_t pixel, uint64_t frame) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = (uint32_t)frame; ctr[2] = (uint32_t)pixel; ctr[3] = (uint32_t)(pixel >> 32) ^ (uint32_t)(frame >> 32);
    have = 0;
  }
  __device__ uint32_t next() {
    if (have == 0) {
      uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
      uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
      for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
      out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
      ctr[0]++;
      have = 4;
    }
    return out[--have];
  }
  __device__ float uniform() { return ((float)next() + 0.5f) * 2.3283064365386963e-10f; } // (0,1)
  __device__ float normal() {
    const float u1 = uniform(), u2 = uniform();
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
};

// Marsaglia-Tsang Gamma(shape, scale), same scheme as camera.cu:21-40.
__device__ float gamma_mt(float shape, float scale, Philox &g) {
  float boost = 1.0f;
  float alpha = shape;
  if (alpha < 1.0f) {
    boost = powf(g.uniform(), 1.0f / alpha);
    alpha += 1.0f;
  }
  const float d = alpha - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
  for (int it = 0; it < 64; ++it) {
    const float z = g.normal(), u = g.uniform();
    const float t = 1.0f + c * z;
    const float v = t * t * t;
    if (t > 0.0f && logf(u) < 0.5f * z * z + d - d * v + d * logf(v)) return d * v * scale * boost;
  }
  return d * scale * boost;
}

struct Src {
  const uint8_t *u8;
  const float *rgba;
  const float *mapx, *mapy;
};

template <bool RGBA>
__device__ __forceinline__ int fetch(const FrontParams &p, const Src &s, int n, int y, int x,
                                     int which) {
  if (x < 0 || x >= p.cols || y < 0 || y >= p.rows) return 0; // zero padding (csct.cu:40-42)
  int fx = x + p.bx, fy = y + p.by;
  if (s.mapx) {
    const size_t mp = (size_t)fy * p.fcols + fx;
    float sx = roundf(__ldg(s.mapx + mp)), sy = roundf(__ldg(s.mapy + mp));
    sx = fminf(fmaxf(sx, 0.0f), (float)(p.fcols - 1));
    sy = fminf(fmaxf(sy, 0.0f), (float)(p.frows - 1));
    fx = (int)sx; fy = (int)sy;
  }
  const size_t sp = ((size_t)n * p.frows + fy) * p.fcols + fx;
  int v;
  if (RGBA) {
    const int t = (int)(__ldg(s.rgba + 4 * sp) * 255); // truncation, core.cu:51
    v = min(max(t, 0), 255);
  } else {
    v = __ldg(s.u8 + sp);
  }
  if (p.speckle_shape > 0.0f) { // camera.cu:66-74
    Philox g(p.seed + (uint64_t)which, (uint64_t)sp, p.frame);
    const float r = roundf((float)v * gamma_mt(p.speckle_shape, p.speckle_scale, g) + p.gaussian_mu +
                           p.gaussian_sigma * g.normal());
    v = min(max((int)r, 0), 255);
  }
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

cudaError_t launch_front(const FrontParams &p, cudaStream_t st) {
  if (p.N > 65535) return cudaErrorInvalidValue;
  const dim3 grid((p.cols + FT - 1) / FT, (p.rows + FT - 1) / FT, p.N);
  const dim3 block(FT, FTY);
  const size_t smem = 2 * (size_t)(FT + p.cw - 1) * (FT + p.ch - 1);
  if (p.left_rgba) front_kernel<true><<<grid, block, smem, st>>>(p);
  else front_kernel<false><<<grid, block, smem, st>>>(p);
  return cudaGetLastError();
}

} // namespace ssb
