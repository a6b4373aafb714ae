// generic.cu -- completeness fallback for configurations outside the packed-u16 fast path of
// aggr.cu (max_disp not a multiple of the lane width, or penalties so large that
// 4*(Cmax+P2) >= 65536).  Same semantics (3rd_party/simsense/src/aggr.cu:29-230,
// src/wta.cu:170-214), int32 math, any D <= 1024.  Still GPU code -- there is no CPU fallback --
// but organised for generality, not speed: one block per path, one thread per disparity.
#include "common.cuh"
#include "kernels.h"
#include <limits.h>

namespace ssb {

__device__ __forceinline__ int block_min(int v, int *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = __reduce_min_sync(FULL, v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int r = lane < nw ? red[lane] : INT_MAX;
  r = __reduce_min_sync(FULL, r);
  __syncthreads();
  return r;
}

__global__ void aggr_generic_kernel(const uint16_t *__restrict__ C, uint16_t *__restrict__ L, int N,
                                    int rows, int cols, int D, int vertical, int reverse, int P1,
                                    int P2) {
  __shared__ int prev[1024 + 2];
  __shared__ int red[32];
  const int per_env = vertical ? cols : rows;
  const int n = blockIdx.x / per_env, q = blockIdx.x % per_env;
  const int steps = vertical ? rows : cols;
  const int d = threadIdx.x;
  const size_t env = (size_t)n * rows * cols * D;
  long stride = vertical ? (long)cols * D : (long)D;
  size_t base = env + (vertical ? (size_t)q * D : (size_t)q * cols * D);
  if (reverse) { base += (size_t)(steps - 1) * stride; stride = -stride; }
  int *pv = prev + 1;
  for (int s = 0; s < steps; ++s) {
    const size_t off = base + (long)s * stride + d;
    const int c = d < D ? C[off] : 0;
    int val = c;
    if (s > 0) {
      const int m = block_min(d < D ? pv[d] : INT_MAX, red);
      if (d < D) {
        int best = pv[d];
        if (d > 0) best = min(best, pv[d - 1] + P1);
        if (d < D - 1) best = min(best, pv[d + 1] + P1);
        best = min(best, m + P2);
        val = c + best - m;
      }
      __syncthreads();
    }
    if (d < D) { pv[d] = val; L[off] = (uint16_t)val; }
    __syncthreads();
  }
}

__global__ void blend_generic_kernel(const uint16_t *L0, const uint16_t *L1, const uint16_t *L2,
                                     const uint16_t *L3, uint16_t *LAll, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) LAll[i] = (uint16_t)(((int)L3[i] + (int)L0[i] + (int)L1[i] + (int)L2[i]) / 4);
}

// one warp per pixel
__global__ void wta_generic_kernel(const uint16_t *__restrict__ LAll, float *__restrict__ dispL,
                                   uint16_t *__restrict__ dispR, size_t npix, int cols, int D,
                                   int uniq) {
  const size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31;
  const int x = (int)(pix % cols);
  const uint16_t *v = LAll + pix * D;
  uint32_t lk = 0xffffffffu, rk = 0xffffffffu;
  for (int d = lane; d < D; d += 32) {
    lk = min(lk, ((uint32_t)v[d] << 16) | (uint32_t)d);
    if (x + d < cols) rk = min(rk, ((uint32_t)LAll[(pix + d) * D + d] << 16) | (uint32_t)d);
  }
  lk = __reduce_min_sync(FULL, lk);
  rk = __reduce_min_sync(FULL, rk);
  const int mval = (int)(lk >> 16), dstar = (int)(lk & 0xffffu);
  bool ok = true;
  for (int d = lane; d < D; d += 32)
    ok = ok && ((int)v[d] * (100 - uniq) >= mval * 100 || abs(dstar - d) <= 1);
  ok = __all_sync(FULL, ok);
  if (lane == 0) {
    float disp = (float)dstar;
    if (!ok) disp = -1.0f;
    else if (dstar != 0 && dstar != D - 1) {
      const int y0 = v[dstar - 1], y2 = v[dstar + 1];
      const float sub = (float)((1.0 * (double)(y2 - y0)) / (2.0 * (double)(y0 - 2 * mval + y2)));
      disp = (float)dstar - sub;
    }
    dispL[pix] = disp;
    dispR[pix] = (uint16_t)(rk & 0xffffu);
  }
}

// b.dbgL0 / b.dbgL3 / b.dbgLAll MUST be valid volumes here (the engine allocates them on demand).
cudaError_t launch_aggr_wta_generic(const AggrBuffers &b, uint16_t *, int N, int rows, int cols,
                                    int D, int P1, int P2, int uniq, cudaStream_t st) {
  const int threads = ((D + 31) / 32) * 32;
  if (threads > 1024 || !b.dbgL0 || !b.dbgL3 || !b.dbgLAll) return cudaErrorInvalidValue;
  aggr_generic_kernel<<<(unsigned)(N * rows), threads, 0, st>>>(b.C, b.dbgL0, N, rows, cols, D, 0, 0, P1, P2);
  aggr_generic_kernel<<<(unsigned)(N * rows), threads, 0, st>>>(b.C, b.L1, N, rows, cols, D, 0, 1, P1, P2);
  aggr_generic_kernel<<<(unsigned)(N * cols), threads, 0, st>>>(b.C, b.L2, N, rows, cols, D, 1, 0, P1, P2);
  aggr_generic_kernel<<<(unsigned)(N * cols), threads, 0, st>>>(b.C, b.dbgL3, N, rows, cols, D, 1, 1, P1, P2);
  const size_t total = (size_t)N * rows * cols * D;
  blend_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(b.dbgL0, b.L1, b.L2, b.dbgL3, b.dbgLAll, total);
  const size_t npix = (size_t)N * rows * cols;
  wta_generic_kernel<<<(unsigned)((npix + 7) / 8), 256, 0, st>>>(b.dbgLAll, b.dispL, b.dispR, npix, cols, D, uniq);
  return cudaGetLastError();
}

} // namespace ssb
