// post.cu -- disparity post-processing and depth output for sm_100a.
//
// Replaces lrConsistencyCheck (3rd_party/simsense/src/lrcheck.cu:21-32), medianFilter
// (src/filter.cu:98-117), pasteSubArea (src/camera.cu:141-158), disp2Depth (:160-168),
// initRgbDepth / depthRegistration / depthDilation / correctDepthRange (:170-240) and the two
// point-cloud kernels (:242-286).  The reference runs these as 8 launches with device-wide
// syncs; here: LR check + median + ROI paste in one tile kernel (median by a register sorting
// network instead of a per-thread selection sort in shared memory), disparity->depth +
// registration splat in one, dilation + range clamp in one.
#include "common.cuh"
#include "kernels.h"

namespace ssb {

// ------------------------------------------------------------------ LR check + median + paste
template <int N> __device__ __forceinline__ void bitonic_sort(float (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
          if ((i & k) == 0) { a[i] = lo; a[l] = hi; }
          else { a[i] = hi; a[l] = lo; }
        }
      }
    }
  }
}

// Median of 9 by the 19-exchange selection network (the result is an order statistic, so any
// correct network reproduces filter.cu:21-44 bit for bit).
__device__ __forceinline__ void cswap(float &a, float &b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo; b = hi;
}
template <int NP> __device__ __forceinline__ float median9(float (&p)[NP]) {
  cswap(p[1], p[2]); cswap(p[4], p[5]); cswap(p[7], p[8]);
  cswap(p[0], p[1]); cswap(p[3], p[4]); cswap(p[6], p[7]);
  cswap(p[1], p[2]); cswap(p[4], p[5]); cswap(p[7], p[8]);
  cswap(p[0], p[3]); cswap(p[5], p[8]); cswap(p[4], p[7]);
  cswap(p[3], p[6]); cswap(p[1], p[4]); cswap(p[2], p[5]);
  cswap(p[4], p[7]); cswap(p[4], p[2]); cswap(p[6], p[4]);
  cswap(p[4], p[2]);
  return p[4];
}

constexpr int PT_X = 32, PT_Y = 8;

__device__ __forceinline__ float atomic_min_float(float *addr, float value) { // camera.cu:42-47
  return (value >= 0) ? __int_as_float(atomicMin((int *)addr, __float_as_int(value)))
                      : __uint_as_float(atomicMax((unsigned int *)addr, __float_as_uint(value)));
}

// disparity -> depth (camera.cu:160-168) and, with registration, the forward splat into the RGB
// frame (camera.cu:179-196); same expression shapes as the reference so nvcc contracts identically.
// (u, v) = full-image pixel coordinates of `pos`
__device__ __forceinline__ void depth_and_splat(const PostParams &p, size_t n, size_t pos, int u, int v, float d) {
  const size_t idx = n * (size_t)p.frows * p.fcols + pos;
  const float z = (d <= 0) ? 0 : p.focal * p.baseline / d;
  p.depth[idx] = z;
  if (p.registration) {
    float a1, a2, a3;
    if (p.a1) {
      a1 = p.a1[pos]; a2 = p.a2[pos]; a3 = p.a3[pos];
    } else { // the float64 products and sums of simsense_component.py:308-325, rounded to float32 like the plane upload
      const double du = (double)u, dv = (double)v;
      a1 = (float)__dadd_rn(__dadd_rn(__dmul_rn(p.reg_m[0], du), __dmul_rn(p.reg_m[1], dv)), p.reg_m[2]);
      a2 = (float)__dadd_rn(__dadd_rn(__dmul_rn(p.reg_m[3], du), __dmul_rn(p.reg_m[4], dv)), p.reg_m[5]);
      a3 = (float)__dadd_rn(__dadd_rn(__dmul_rn(p.reg_m[6], du), __dmul_rn(p.reg_m[7], dv)), p.reg_m[8]);
    }
    const float zRgb = a3 * z + p.b3;
    const int x = (int)roundf((a1 * z + p.b1) / zRgb);
    const int y = (int)roundf((a2 * z + p.b2) / zRgb);
    if (zRgb > 0 && x >= 0 && x < p.rgb_cols && y >= 0 && y < p.rgb_rows)
      atomic_min_float(p.canvas + (n * p.rgb_rows + y) * p.rgb_cols + x, zRgb);
  } else {
    p.out[idx] = (z < p.min_depth || z >= p.max_depth) ? 0.0f : z; // camera.cu:237-239
  }
}

// Gate of a column band: ONE warp waits (with back-off) until the final aggregation pass has bumped
// its progress counter to `target`; the band's kernels follow in stream order.  A single waiting warp
// cannot keep the pass's own blocks off an SM, whatever order the hardware starts the kernels in
// (band kernels whose every block polled could: 8 such blocks fill the thread slots of an SM).
__global__ void gate_kernel(const uint32_t *ctr, uint32_t target) {
  if (threadIdx.x == 0) {
    unsigned v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      __nanosleep(500);
    }
  }
}
cudaError_t launch_gate(const uint32_t *ctr, uint32_t target, cudaStream_t st) {
  gate_kernel<<<1, 32, 0, st>>>(ctr, target);
  return cudaGetLastError();
}

// FUSE: no ROI -> the matched image is the full image, so depth + splat run in the same thread.
template <int K, bool FUSE>
__global__ void __launch_bounds__(PT_X *PT_Y) lr_median_kernel(const PostParams p) {
  constexpr int H = K / 2;
  constexpr int WC = PT_X + 2 * H, WR = PT_Y + 2 * H;
  __shared__ float tile[WR][WC];
  const int n = blockIdx.z;
  const int x0 = p.xa + blockIdx.x * PT_X, y0 = blockIdx.y * PT_Y; // columns [xa, xb) of every row
  const size_t img = (size_t)n * p.rows * p.cols;
  const int tid = threadIdx.y * PT_X + threadIdx.x;
  // Tile load in straight-line phases (left disparities, then the dependent right-disparity
  // gathers) so that the loads of all rounds are in flight together.
  constexpr int NT = PT_X * PT_Y, NEL = WR * WC, NRND = (NEL + NT - 1) / NT;
  float v[NRND];
  size_t pos[NRND];
  int xs[NRND];
  bool inside[NRND];
  // (the last round covers only NEL - (NRND-1)*NT elements -- 84 of 256 threads for the 3x3 median: warps that have no
  //  element in a round skip it, a warp-uniform test)
  const int wbase = tid & ~31;
#pragma unroll
  for (int k = 0; k < NRND; ++k) {
    inside[k] = false; xs[k] = 0; pos[k] = img; v[k] = -1.0f;
    if (k * NT + wbase >= NEL) continue;
    const int i = min(k * NT + tid, NEL - 1);
    const int wy = i / WC, wx = i - wy * WC;
    const int y = y0 + wy - H, x = x0 + wx - H;
    inside[k] = x >= 0 && x < p.cols && y >= 0 && y < p.rows;
    xs[k] = min(max(x, 0), p.cols - 1);
    pos[k] = img + (size_t)min(max(y, 0), p.rows - 1) * p.cols + xs[k];
    v[k] = p.dispL[pos[k]];
  }
  if (p.lr_max_diff != 255) { // lrcheck.cu:28-31
    int ld[NRND];
    uint16_t dr[NRND];
#pragma unroll
    for (int k = 0; k < NRND; ++k) {
      ld[k] = 0; dr[k] = 0;
      if (k * NT + wbase >= NEL) continue;
      ld[k] = (int)roundf(v[k]);
      dr[k] = p.dispR[pos[k] - min(max(ld[k], 0), xs[k])];
    }
#pragma unroll
    for (int k = 0; k < NRND; ++k)
      if (ld[k] < 0 || xs[k] - ld[k] < 0 || abs(ld[k] - (int)dr[k]) > p.lr_max_diff) v[k] = -1.0f;
  }
#pragma unroll
  for (int k = 0; k < NRND; ++k) {
    const int i = k * NT + tid;
    if (i < NEL) {
      const int wy = i / WC, wx = i - wy * WC;
      if (p.disp_lr && inside[k] && wy >= H && wy < WR - H && wx >= H && wx < WC - H) p.disp_lr[pos[k]] = v[k];
      tile[wy][wx] = inside[k] ? v[k] : -1.0f;
    }
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= p.xb || y >= p.rows) return;
  float out = tile[threadIdx.y + H][threadIdx.x + H];
  if (K > 1 && x >= H && y >= H && x < p.cols - H && y < p.rows - H) { // filter.cu:107
    constexpr int NN = K * K;
    constexpr int NP = NN <= 16 ? 16 : (NN <= 32 ? 32 : 64);
    float a[NP];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) a[j * K + i] = tile[threadIdx.y + j][threadIdx.x + i];
    if constexpr (K == 3) {
      out = median9(a);
    } else {
#pragma unroll
      for (int i = NN; i < NP; ++i) a[i] = __int_as_float(0x7f800000);
      bitonic_sort<NP>(a);
      out = a[NN / 2];
    }
  }
  p.disp_med[img + (size_t)y * p.cols + x] = out;
  if (FUSE) depth_and_splat(p, (size_t)n, (size_t)y * p.cols + x, x, y, out);
  else if (p.bbox) p.disp_full[((size_t)n * p.frows + (y + p.by)) * p.fcols + x + p.bx] = out;
}

// ------------------------------------------------------------------ depth + registration splat
__global__ void fill_kernel(float *dst, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

__global__ void __launch_bounds__(256) depth_splat_kernel(const PostParams p) { // ROI mode only
  const size_t fsz = (size_t)p.frows * p.fcols;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= fsz * p.N) return;
  const size_t n = idx / fsz;
  const size_t pos = idx - n * fsz;
  const int v = (int)(pos / p.fcols);
  depth_and_splat(p, n, pos, (int)(pos - (size_t)v * p.fcols), v, p.disp_full[idx]);
}

// Dilation with snapshot semantics + range clamp, canvas -> out.
// camera.cu:207-227: a pixel < maxDepth lowers its left, top and top-left neighbours, i.e.
// out(x,y) = min over the 2x2 block (x..x+1, y..y+1) of the inputs that are < maxDepth.
// 2-D grid (no index divisions), 4 pixels per thread.
__global__ void __launch_bounds__(128) dilate_range_kernel(const PostParams p) {
  const int x4 = p.ua + (blockIdx.x * blockDim.x + threadIdx.x) * 4; // columns [ua, ub): ua % 4 == 0
  const int y = blockIdx.y;
  if (x4 >= p.ub) return;
  const size_t row = ((size_t)blockIdx.z * p.rgb_rows + y) * p.rgb_cols;
  const float *r0 = p.canvas + row + x4;
  const bool yb = y + 1 < p.rgb_rows;
  const float *r1 = r0 + (yb ? p.rgb_cols : 0);
  float a[5], b[5];
  const bool vec = (p.rgb_cols & 3) == 0; // rows are 16-byte aligned and x4+3 < cols
  if (vec) {
    const float4 va = *reinterpret_cast<const float4 *>(r0), vb = *reinterpret_cast<const float4 *>(r1);
    a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
    b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool in = x4 + k < p.rgb_cols;
      a[k] = in ? r0[k] : p.max_depth;
      b[k] = in ? r1[k] : p.max_depth;
    }
  }
  const bool xr = x4 + 4 < p.rgb_cols;
  a[4] = xr ? r0[4] : p.max_depth;
  b[4] = xr ? r1[4] : p.max_depth;
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float m = a[k];
    if (p.dilation) {
      if (a[k + 1] < p.max_depth) m = fminf(m, a[k + 1]);
      if (yb && b[k] < p.max_depth) m = fminf(m, b[k]);
      if (yb && b[k + 1] < p.max_depth) m = fminf(m, b[k + 1]);
    }
    o[k] = (m < p.min_depth || m >= p.max_depth) ? 0.0f : m; // camera.cu:237-239
  }
  float *dst = p.out + row + x4;
  if (vec) {
    *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x4 + k < p.rgb_cols) dst[k] = o[k];
  }
}

cudaError_t launch_post(const PostParams &p_in, cudaStream_t st, int *launches) {
  PostParams p = p_in;
  if (p.N > 65535) return cudaErrorInvalidValue;
  const bool band = p.xb > 0 || p.ub > 0; // a column band of a frame whose setup launches already ran
  if (!band) { p.xa = 0; p.xb = p.cols; p.ua = 0; p.ub = p.rgb_cols; }
  if (band && (p.bbox || (p.xa & 31) || (p.ua & 3) || p.xb > p.cols || p.ub > p.rgb_cols)) return cudaErrorInvalidValue;
  int nl = 0;
  cudaError_t err;
  const size_t fsz = (size_t)p.frows * p.fcols * p.N;
  const size_t rsz = (size_t)p.rgb_rows * p.rgb_cols * p.N;
  if (p.bbox) { // outside-ROI disparity is defined as 0 (reference leaves it uninitialised)
    if ((err = cudaMemsetAsync(p.disp_full, 0, fsz * sizeof(float), st)) != cudaSuccess) return err;
  }
  if (p.registration && !p.canvas_prefilled && !band) {
    fill_kernel<<<(unsigned)min((rsz + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(p.canvas, rsz, p.max_depth);
    ++nl;
  }
  const dim3 grid((p.xb - p.xa + PT_X - 1) / PT_X, (p.rows + PT_Y - 1) / PT_Y, p.N);
  const dim3 block(PT_X, PT_Y);
#define SSB_MED(KK)                                                                               \
  case KK:                                                                                        \
    if (p.bbox) lr_median_kernel<KK, false><<<grid, block, 0, st>>>(p);                           \
    else lr_median_kernel<KK, true><<<grid, block, 0, st>>>(p);                                   \
    break;
  if (p.xb > p.xa) {
    switch (p.mf_size) {
      SSB_MED(1) SSB_MED(3) SSB_MED(5) SSB_MED(7)
    default: return cudaErrorInvalidValue;
    }
    ++nl;
  }
#undef SSB_MED
  if (p.bbox) {
    depth_splat_kernel<<<(unsigned)((fsz + 255) / 256), 256, 0, st>>>(p);
    ++nl;
  }
  if (p.registration && p.ub > p.ua) {
    const int span = p.ub - p.ua; // 4 pixels per thread; narrow images / bands get narrower blocks instead of idle warps
    const int threads = span <= 128 ? 32 : (span <= 256 ? 64 : 128);
    const dim3 dg((unsigned)((span + 4 * threads - 1) / (4 * threads)), (unsigned)p.rgb_rows, (unsigned)p.N);
    dilate_range_kernel<<<dg, threads, 0, st>>>(p);
    ++nl;
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------ point clouds
template <bool RGB>
__global__ void __launch_bounds__(256)
point_cloud_kernel(const float *__restrict__ depth, const float *__restrict__ rgba,
                   float *__restrict__ pc, size_t total, int rows, int cols, float fx, float fy,
                   float s, float cx, float cy) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t pos = idx % ((size_t)rows * cols);
  const int u = (int)(pos % cols), v = (int)(pos / cols);
  const float z = depth[idx]; // camera.cu:250-260
  const float x = z * ((u - cx) / fx + s * (cy - v) / (fx * fy));
  const float y = z * (v - cy) / fy;
  if (RGB) {
    const float4 c = __ldg(reinterpret_cast<const float4 *>(rgba) + idx);
    float2 *o = reinterpret_cast<float2 *>(pc + 6 * idx);
    o[0] = make_float2(x, y); o[1] = make_float2(z, c.x); o[2] = make_float2(c.y, c.z);
  } else {
    pc[3 * idx] = x; pc[3 * idx + 1] = y; pc[3 * idx + 2] = z;
  }
}

cudaError_t launch_point_cloud(const float *depth, const float *rgba, float *pc, int N, int rows,
                               int cols, float fx, float fy, float skew, float cx, float cy,
                               cudaStream_t st) {
  const size_t total = (size_t)N * rows * cols;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (rgba) point_cloud_kernel<true><<<blocks, 256, 0, st>>>(depth, rgba, pc, total, rows, cols, fx, fy, skew, cx, cy);
  else point_cloud_kernel<false><<<blocks, 256, 0, st>>>(depth, nullptr, pc, total, rows, cols, fx, fy, skew, cx, cy);
  return cudaGetLastError();
}

} // namespace ssb
