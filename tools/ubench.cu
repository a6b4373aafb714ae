// tools/ubench.cu -- instruction throughput / dependent-issue latency probes for the integer ops the
// stereo kernels live on (POPC, LOP3, IMAD, PRMT, DPX min/add, REDUX, SHFL).  Not product code:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define FULL 0xffffffffu

enum Op { POPC, LOP, IMAD, PRMT, IADD, VIADDMIN, VMIN3, REDUX, SHFL, POPC_LOP, NOPS };
static const char *names[] = {"popc", "lop3(xor)", "imad", "prmt", "iadd3", "viaddmin_u16x2", "vimin3_u16x2", "redux.min", "shfl.up", "popc+xor pair"};

template <int OP> __device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t k) {
  if (OP == POPC) return __popc(x) + k;              // the +k keeps the chain data dependent (1 extra IADD folded below)
  if (OP == LOP) return x ^ k;
  if (OP == IMAD) return x * k + k;
  if (OP == PRMT) return __byte_perm(x, k, 0x1230);
  if (OP == IADD) return x + k;
  if (OP == VIADDMIN) return __viaddmin_u16x2(x, k, k);
  if (OP == VMIN3) return __vimin3_u16x2(x, k, k ^ 0x10001u);
  if (OP == REDUX) return __reduce_min_sync(FULL, x) + k;
  if (OP == SHFL) return __shfl_up_sync(FULL, x, 1) + k;
  if (OP == POPC_LOP) return __popc(x ^ k) + k;
  return x;
}

// ILP independent chains per thread, ITER iterations
template <int OP, int ILP> __global__ void probe(uint32_t *out, uint32_t k, int iters, long long *cyc) {
  uint32_t v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 7 + i + k;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = apply<OP>(v[i], k);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(uint32_t *out, long long *cyc) {
  const int iters = 4096;
  long long h = 0;
  // latency: one warp, one chain
  probe<OP, 1><<<1, 32>>>(out, 3, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double lat = (double)h / iters;
  // throughput: 148 blocks x 512 threads (4 warps / SMSP), 8 chains each
  probe<OP, 8><<<148, 512>>>(out, 3, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_sm_clk = (double)iters * 8 * 512 / (double)h; // thread-ops per clock per SM
  // one warp per SMSP, 8 chains
  probe<OP, 8><<<148, 128>>>(out, 3, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_sm_clk1 = (double)iters * 8 * 128 / (double)h;
  printf("%-16s chain step %.1f cyc | %.1f thread-ops/clk/SM (16 warps/SM) | %.1f (4 warps/SM)\n", names[OP], lat, per_sm_clk, per_sm_clk1);
}

// POPC mixed with shared-memory loads / stores: do the xu pipe and the lsu share a dispatch port?
template <int NPOPC, int NLDS, int NSTS, int NALU> __global__ void mix(uint32_t *out, uint32_t k, int iters, long long *cyc) {
  __shared__ uint32_t sm[32 * 16 * 9];
  uint32_t p[8], l[8], a[8];
  uint32_t *mine = sm + (threadIdx.x >> 5) * 32 * 9 + (threadIdx.x & 31);
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = threadIdx.x + i + k; l[i] = 0; a[i] = i; mine[i * 32] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NPOPC) p[i] = __popc(p[i]) + k;
      if (i < NLDS) l[i] += *(volatile uint32_t *)(mine + i * 32);
      if (i < NSTS) *(volatile uint32_t *)(mine + i * 32) = a[i];
      if (i < NALU) a[i] = __byte_perm(a[i], k, 0x1230);
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= p[i] ^ l[i] ^ a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NPOPC, int NLDS, int NSTS, int NALU> void runmix(uint32_t *out, long long *cyc) {
  const int iters = 2048;
  long long h = 0;
  mix<NPOPC, NLDS, NSTS, NALU><<<148, 512>>>(out, 3, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mix popc=%d lds=%d sts=%d alu=%d per iteration: %.1f cycles/iteration/SMSP-warp-set (16 warps/SM => 4 warps per SMSP)\n", NPOPC, NLDS, NSTS, NALU, (double)h / iters);
}

int main() {
  uint32_t *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 8);
  printf("NB: popc/redux/shfl chains include one dependent IADD (subtract the iadd3 chain step)\n");
  run<IADD>(out, cyc);
  run<LOP>(out, cyc);
  run<IMAD>(out, cyc);
  run<PRMT>(out, cyc);
  run<POPC>(out, cyc);
  run<POPC_LOP>(out, cyc);
  run<VIADDMIN>(out, cyc);
  run<VMIN3>(out, cyc);
  run<REDUX>(out, cyc);
  run<SHFL>(out, cyc);
  runmix<8, 0, 0, 0>(out, cyc);
  runmix<0, 8, 0, 0>(out, cyc);
  runmix<0, 0, 8, 0>(out, cyc);
  runmix<0, 0, 0, 8>(out, cyc);
  runmix<8, 8, 0, 0>(out, cyc);
  runmix<8, 0, 8, 0>(out, cyc);
  runmix<8, 0, 0, 8>(out, cyc);
  runmix<8, 8, 8, 8>(out, cyc);
  runmix<4, 8, 8, 8>(out, cyc);
  runmix<2, 8, 8, 8>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
