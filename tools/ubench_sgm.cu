// tools/ubench_sgm.cu -- how many cycles does one SGM path step cost a warp, and what does interleaving several
// independent paths in ONE warp (ILP) buy?  Not product code (sgm_step is the one of sapien_b200/csrc/aggr.cu).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_sgm tools/ubench_sgm.cu && gpurun_out/ubench_sgm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu

template <int NR>
__device__ __forceinline__ void sgm_step(uint32_t (&L)[NR], const uint32_t (&c)[NR], uint32_t P1P1, uint32_t P2P2, uint32_t selUp, uint32_t selDn) {
  uint32_t t = L[0];
#pragma unroll
  for (int j = 1; j < NR; ++j) t = __vminu2(t, L[j]);
  t = __vminu2(t, __byte_perm(t, t, 0x1032));
  const uint32_t mm = __reduce_min_sync(FULL, t);
  const uint32_t mP2 = mm + P2P2;
  const uint32_t up = __shfl_up_sync(FULL, L[NR - 1], 1);
  const uint32_t dn = __shfl_down_sync(FULL, L[0], 1);
  uint32_t nl[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const uint32_t lm1 = (j == 0) ? __byte_perm(up, L[0], selUp) : __funnelshift_l(L[j - 1], L[j], 16);
    const uint32_t lp1 = (j == NR - 1) ? __byte_perm(L[NR - 1], dn, selDn) : __funnelshift_r(L[j], L[j + 1], 16);
    uint32_t v = __viaddmin_u16x2(lm1, P1P1, L[j]);
    v = __viaddmin_u16x2(lp1, P1P1, v);
    v = __vminu2(v, mP2);
    nl[j] = v - mm + c[j];
  }
#pragma unroll
  for (int j = 0; j < NR; ++j) L[j] = nl[j];
}

template <int NR, int ILP> __global__ void probe(uint32_t *out, int steps, long long *cyc, uint32_t P1, uint32_t P2) {
  extern __shared__ uint32_t sm[]; // [warps][ILP][64 steps][32 lanes][NR]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *mine = sm + (size_t)warp * ILP * 64 * 32 * NR;
  for (int i = lane; i < ILP * 64 * 32 * NR; i += 32) mine[i] = (i * 2654435761u) & 0x03ff03ffu;
  __syncthreads();
  uint32_t L[ILP][NR];
#pragma unroll
  for (int p = 0; p < ILP; ++p)
#pragma unroll
    for (int r = 0; r < NR; ++r) L[p][r] = 0;
  const uint32_t selUp = lane == 0 ? 0x5454u : 0x5432u, selDn = lane == 31 ? 0x3232u : 0x5432u;
  const long long t0 = clock64();
  for (int s = 0; s < steps; s += 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int p = 0; p < ILP; ++p) {
        uint32_t c[NR];
        const uint32_t *src = mine + ((size_t)(p * 64 + ((s + k) & 63)) * 32 + lane) * NR;
#pragma unroll
        for (int r = 0; r < NR; ++r) c[r] = src[r];
        sgm_step<NR>(L[p], c, P1, P2, selUp, selDn);
        // the pass also stores the new L every step
        if (NR == 2) *reinterpret_cast<uint2 *>(out + ((size_t)(blockIdx.x * blockDim.x + threadIdx.x) * ILP + p) * NR) = make_uint2(L[p][0], L[p][1]);
        else out[((size_t)(blockIdx.x * blockDim.x + threadIdx.x) * ILP + p) * NR] = L[p][0];
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NR, int ILP> void run(uint32_t *out, long long *cyc, int warps) {
  const int steps = 4096;
  long long h = 0;
  const size_t smem = (size_t)warps * ILP * 64 * 32 * NR * 4;
  cudaFuncSetAttribute(probe<NR, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<NR, ILP><<<148, 32 * warps, smem>>>(out, steps, cyc, 0x01880188u, 0x06200620u);
  cudaDeviceSynchronize();
  probe<NR, ILP><<<148, 32 * warps, smem>>>(out, steps, cyc, 0x01880188u, 0x06200620u);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("NR=%d  paths per warp (ILP)=%d  warps/SM=%d : %.1f cycles per step of one warp = %.1f cycles per path-step; SM throughput %.3f path-steps/clk\n",
         NR, ILP, warps, (double)h / steps, (double)h / steps / ILP, (double)warps * ILP * steps / (double)h);
}

int main() {
  uint32_t *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4 * 4 * 4); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 5, 8}) { run<2, 1>(out, cyc, w); run<2, 2>(out, cyc, w); if (w <= 4) run<2, 3>(out, cyc, w); }
  for (int w : {1, 4, 8}) { run<1, 1>(out, cyc, w); run<1, 2>(out, cyc, w); run<1, 4>(out, cyc, w); }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
