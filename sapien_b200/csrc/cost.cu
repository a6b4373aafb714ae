// cost.cu -- Hamming cost volume fused with the separable block sum, for sm_100a.
//
// Replaces hammingCost (3rd_party/simsense/src/cost.cu:21-47) + boxFilterHorizontal /
// boxFilterVertical (src/filter.cu:49-96).  The reference writes the raw Hamming volume, re-reads
// it for a serial horizontal running sum through global memory, and again for the vertical one
// (3 volumes written, 2 read).  Here one kernel writes the final volume once:
//
//   C(y,x,d) = sum_{j=-hh..hh} sum_{i=-hw..hw} ham(clamp(y+j), clamp(x+i), d),
//   ham(y,x,d) = popc(cL(y,x) ^ cR(y, max(x-d,0)))                       
//
// Mapping: a thread owns a disparity PAIR (packed u16x2) for a strip of TX output columns and
// marches down a band of rows.  Per input row it computes TX+BW-1 Hamming pairs from census codes
// staged in shared memory (1 LDS + 2 POPC per column), a sliding BW-wide horizontal sum in
// registers, and a BH-deep vertical running sum whose leaving row comes from a thread-private
// shared-memory ring.  ~ (TX+BW-1)/TX * (RY+BH-1)/RY POPC per output instead of BW*BH; stores are
// 4 B per lane, 128 B per warp, and every volume byte is written exactly once.
#include "common.cuh"
#include "kernels.h"

namespace ssb {

constexpr int COST_RY = 48; // output rows per band

template <int BW, int BH, int TX, int NS, int TD>
__global__ void __launch_bounds__(TD *NS)
cost_kernel(const uint32_t *__restrict__ cL, const uint32_t *__restrict__ cR,
            uint16_t *__restrict__ C, int rows, int cols, int D, int nchunks) {
  constexpr int HW = BW / 2, HH = BH / 2;
  constexpr int NH = TX + BW - 1;       // hamming columns per strip
  constexpr int NXW = NS * TX + BW - 1; // left census codes staged per block
  constexpr int DC = 2 * TD;            // disparities per chunk
  constexpr int NRC = NXW + DC;         // right census codes staged per block
  constexpr int SLOT = NS * TX * TD;    // ring words per input row
  __shared__ uint32_t sL[2][NXW];
  __shared__ uint32_t sR[2][NRC];
  extern __shared__ uint32_t ring[]; // [BH][NS][TX][TD], thread-private entries

  const int td = threadIdx.x; // disparity pair inside the chunk
  const int strip = threadIdx.y;
  const int tid = strip * TD + td;
  constexpr int nthreads = TD * NS;
  const int chunk = blockIdx.x % nchunks;
  const int xblk = blockIdx.x / nchunks;
  const int n = blockIdx.z;
  const int dbase = chunk * DC;
  const int d_lo = dbase + 2 * td;       // my disparities: d_lo, d_lo+1
  const int xs = xblk * (NS * TX) - HW;  // image column of staged index 0
  const int y_begin = blockIdx.y * COST_RY;
  const int y_end = min(rows, y_begin + COST_RY);
  const uint32_t *imL = cL + (size_t)n * rows * cols;
  const uint32_t *imR = cR + (size_t)n * rows * cols;
  uint16_t *outC = C + (size_t)n * rows * cols * D;
  const int imax = cols - 1 - xs; // staged index of the last image column (replicate border)
  const bool even_d = (D & 1) == 0;

  auto stage = [&](int buf, int yin) {
    const int yc = min(max(yin, 0), rows - 1);
    const uint32_t *l = imL + (size_t)yc * cols;
    const uint32_t *r = imR + (size_t)yc * cols;
    for (int i = tid; i < NXW; i += nthreads) sL[buf][i] = __ldg(l + min(max(xs + i, 0), cols - 1));
    // sR[j] holds cR(y, max(xs + j - (DC-1) - dbase, 0)); column i / local disparity dl -> j = i-dl+DC-1
    for (int j = tid; j < NRC; j += nthreads)
      sR[buf][j] = __ldg(r + min(max(xs + j - (DC - 1) - dbase, 0), cols - 1));
  };

  uint32_t vacc[TX];
#pragma unroll
  for (int x = 0; x < TX; ++x) vacc[x] = 0;
  uint32_t *myring = ring + (size_t)strip * TX * TD + td;
  const int ib = strip * TX; // staged index of my first hamming column

  // input rows y_begin-HH .. y_end-1+HH (clamped); output row y is complete after input y+HH
  const int nin = (y_end - y_begin) + BH - 1;
  stage(0, y_begin - HH);
  __syncthreads();
  for (int it = 0; it < nin; ++it) {
    const int buf = it & 1;
    if (it + 1 < nin) stage(buf ^ 1, y_begin - HH + it + 1);
    // ---- Hamming pairs of my strip ---------------------------------------------------------
    const uint32_t *pl = sL[buf];
    const uint32_t *pr = sR[buf] + (DC - 1) - 2 * td;
    uint32_t h[NH];
    {
      uint32_t r0 = pr[ib - 1], hcur = 0;
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        if (ib + i <= imax) {
          const uint32_t r1 = r0; // code for d_lo+1 at this column == code for d_lo one column left
          r0 = pr[ib + i];
          const uint32_t a = pl[ib + i];
          hcur = (uint32_t)__popc(a ^ r0) | ((uint32_t)__popc(a ^ r1) << 16);
        }
        h[i] = hcur;
      }
    }
    // ---- sliding BW-sum along x, BH-deep running sum along y -------------------------------
    uint32_t *rs_w = myring + (size_t)(it % BH) * SLOT;
    const uint32_t *rs_r = myring + (size_t)((it + 1) % BH) * SLOT;
    const int y = y_begin + it - (BH - 1);
    const bool emit = it >= BH - 1;
    uint32_t hs = 0;
#pragma unroll
    for (int i = 0; i < BW - 1; ++i) hs += h[i];
#pragma unroll
    for (int x = 0; x < TX; ++x) {
      hs += h[x + BW - 1];
      vacc[x] += hs;
      rs_w[x * TD] = hs;
      hs -= h[x];
      if (emit) {
        const int xo = xs + HW + ib + x;
        if (xo < cols && d_lo < D) {
          uint16_t *dst = outC + ((size_t)y * cols + xo) * D + d_lo;
          if (even_d) {
            *reinterpret_cast<uint32_t *>(dst) = vacc[x];
          } else {
            dst[0] = (uint16_t)(vacc[x] & 0xffffu);
            if (d_lo + 1 < D) dst[1] = (uint16_t)(vacc[x] >> 16);
          }
        }
        vacc[x] -= rs_r[x * TD]; // the row that leaves the window before the next input
      }
    }
    __syncthreads();
  }
}

// Any block size: direct evaluation (bw*bh POPC per output).  Only used for block sizes that have
// no specialisation above.
__global__ void cost_generic_kernel(const uint32_t *__restrict__ cL, const uint32_t *__restrict__ cR,
                                    uint16_t *__restrict__ C, int rows, int cols, int D, int bw,
                                    int bh, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int d = (int)(idx % D);
  const size_t pix = idx / D;
  const int x = (int)(pix % cols);
  const size_t ny = pix / cols;
  const int y = (int)(ny % rows);
  const size_t n = ny / rows;
  const uint32_t *l = cL + n * rows * cols;
  const uint32_t *r = cR + n * rows * cols;
  const int hw = bw / 2, hh = bh / 2;
  unsigned acc = 0;
  for (int j = -hh; j <= hh; ++j) {
    const int yc = min(max(y + j, 0), rows - 1);
    for (int i = -hw; i <= hw; ++i) {
      const int xc = min(max(x + i, 0), cols - 1);
      acc += __popc(l[(size_t)yc * cols + xc] ^ r[(size_t)yc * cols + max(xc - d, 0)]);
    }
  }
  C[idx] = (uint16_t)acc;
}

template <int BW, int BH>
static cudaError_t launch_fast(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                               int cols, int D, cudaStream_t st) {
  constexpr int TX = 16;
  if (D <= 64) {
    constexpr int TD = 32, NS = 4;
    auto k = cost_kernel<BW, BH, TX, NS, TD>;
    const size_t smem = (size_t)BH * NS * TX * TD * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nchunks = (D + 2 * TD - 1) / (2 * TD);
    dim3 grid((unsigned)(((cols + NS * TX - 1) / (NS * TX)) * nchunks),
              (unsigned)((rows + COST_RY - 1) / COST_RY), (unsigned)N);
    k<<<grid, dim3(TD, NS), smem, st>>>(cL, cR, C, rows, cols, D, nchunks);
  } else {
    constexpr int TD = 64, NS = 2;
    auto k = cost_kernel<BW, BH, TX, NS, TD>;
    const size_t smem = (size_t)BH * NS * TX * TD * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nchunks = (D + 2 * TD - 1) / (2 * TD);
    dim3 grid((unsigned)(((cols + NS * TX - 1) / (NS * TX)) * nchunks),
              (unsigned)((rows + COST_RY - 1) / COST_RY), (unsigned)N);
    k<<<grid, dim3(TD, NS), smem, st>>>(cL, cR, C, rows, cols, D, nchunks);
  }
  return cudaGetLastError();
}

cudaError_t launch_cost(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                        int cols, int D, int bw, int bh, cudaStream_t st) {
  if (N > 65535) return cudaErrorInvalidValue;
  if (bw == 7 && bh == 7) return launch_fast<7, 7>(cL, cR, C, N, rows, cols, D, st);
  if (bw == 1 && bh == 1) return launch_fast<1, 1>(cL, cR, C, N, rows, cols, D, st);
  if (bw == 3 && bh == 3) return launch_fast<3, 3>(cL, cR, C, N, rows, cols, D, st);
  if (bw == 5 && bh == 5) return launch_fast<5, 5>(cL, cR, C, N, rows, cols, D, st);
  if (bw == 9 && bh == 9) return launch_fast<9, 9>(cL, cR, C, N, rows, cols, D, st);
  const size_t total = (size_t)N * rows * cols * D;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cost_generic_kernel<<<blocks, 256, 0, st>>>(cL, cR, C, rows, cols, D, bw, bh, total);
  return cudaGetLastError();
}

} // namespace ssb
