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
// marches down a band of rows; a warp = 32 consecutive pairs (64 disparities) of one strip and is
// fully independent (no block barrier).  Per input row it needs TX+BW-1 Hamming pairs (2 XOR +
// 2 POPC + one IMAD that packs them), a sliding BW-wide horizontal sum in registers, and a BH-deep
// vertical running sum whose leaving row comes from a thread-private shared-memory ring (bytes when
// BW * bits <= 255).  (TX+BW-1)/TX * (RY+BH-1)/RY POPC per output instead of BW*BH.
//  * census rows reach shared memory by cp.async, CSTAGES-1 rows ahead, staged per warp with the
//    replicate border / max(x-d,0) clamps already applied.  The right image border costs a warp-uniform
//    select on the last BW/2 columns of the last strip when the strips tile the image; only ragged
//    widths run the EDGE variant (a per-column hold, ~25 % more instructions);
//  * POPC runs on the quarter-rate xu pipe (measured 16 lanes/clk/SM, tools/ubench.cu), the sums on
//    the alu pipe.  The Hamming pairs are produced a few columns AHEAD of the sliding window, one or two
//    per output column, tied to the sliding-sum chain through a run-time zero the compiler cannot see
//    through -- so the two pipes overlap inside every warp instead of taking turns (phase-by-phase code
//    measured xu time + issue time) and only ~20 pairs are live at a time (cost_band_s);
//  * stores are 4 B per lane / 128 B per warp and every volume byte is written once.
#include "common.cuh"
#include "kernels.h"

namespace ssb {

// optional per-block timeline (debug aid, tools/trace_aggr.py): {start ns, end ns, SM id} per block
__device__ unsigned long long *g_cost_trace = nullptr;
__device__ __forceinline__ unsigned long long cost_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int CSTAGES = 4; // census rows in flight per warp (cp.async ring)
template <int BW, int TX, bool MRG = false> struct CostStage { // per-warp staging buffers (words), CSTAGES deep
  static constexpr int NH = TX + BW - 1;
  // MRG (see cost_band_s): one warp serves the upper 32 disparities of TWO adjacent strips
  static constexpr int SLS = (NH + (MRG ? TX : 0) + 3) & ~3;
  static constexpr int SRS = (NH + (MRG ? TX : 0) + 64 + 2 + 3) & ~3;
};

// One band of rows of one strip.  The Hamming pairs are a STREAM: they are produced LEAD = BW-1+KA columns ahead
// of the sliding window and die BW columns later, so only ~BW+LEAD+NH/TX of them are live at any time (the last
// LEAD pairs produced during a row are the first ones of the next row: the stream runs across the row
// boundary).  Round 1 kept the pairs of the whole current and the whole next row in registers (2 * NH of them,
// 254 registers, 8 warps per SM, 628 instructions per row of 32 columns); the stream needs 159 registers (12
// warps per SM: the kernel is latency-bound at 2 warps per scheduler), has no row-end copy, and both running sums
// are one three-input add per column (524 instructions per row):
//   hs(x+1)   = hs(x) + h[x+BW] - h[x]
//   out(y)    = out(y-1) + hs(y) - hs(y-BH)        (ring: BH slots, the slot is read, then overwritten)
__host__ __device__ constexpr bool cost_stream_ok(int BW, int TX, int KA) {
  const int NH = TX + BW - 1, LEAD = BW - 1 + KA;
  if (KA < 1 || LEAD >= NH) return false;
  // item i of the next row must be produced AFTER the last read of item i of this row (the slide of step i)
  for (int x = 0; x < TX; ++x) {
    const int p0 = LEAD + (x * NH) / TX, p1 = LEAD + ((x + 1) * NH) / TX;
    for (int p = p0; p < p1; ++p)
      if (p >= NH && x <= p - NH) return false;
  }
  return true;
}

// MRG (D = 96, strips that tile the image): a block is three warps for two strips -- warps 0 / 1 own disparities
// 0..63 of strip 0 / 1, warp 2 owns disparities 64..95 of BOTH strips (lanes 0-15: strip 0, lanes 16-31: strip 1)
// instead of two half-empty warps.  Its staging buffers cover both strips (the right-image window of the second
// strip is the first one's shifted by TX columns), every lane-dependent quantity (strip, disparity, output
// pointer, border flag) is per lane, and the left codes are a 2-address broadcast.
template <int BW, int BH, int TX, int NS, int TD, bool EDGE, bool PACK8, bool ODD_D, int DT, int KA, bool MRG = false>
__device__ __forceinline__ void cost_band_s(const uint32_t *__restrict__ imL, const uint32_t *__restrict__ imR,
                                            uint16_t *__restrict__ outC, int rows, int cols, int Drt, int dbase,
                                            int xblk, int y_begin, int y_end, uint32_t *ring,
                                            uint32_t *sLb, uint32_t *sRb, uint32_t zmask, bool rb) {
  constexpr int HW = BW / 2, HH = BH / 2;
  constexpr int NH = TX + BW - 1;       // hamming columns per strip
  constexpr int LEAD = BW - 1 + KA;     // pairs produced ahead of the window's left edge
  static_assert(cost_stream_ok(BW, TX, KA), "look-ahead too long for this strip width");
  constexpr int DCW = 64;               // disparities per warp (32 lanes x one pair)
  constexpr int NRC = NH + DCW + 1;     // right census codes staged per warp
  static_assert(!MRG || (NS == 2 && !EDGE && !ODD_D && DT == 96), "merged upper chunks: D = 96, two strips, tiling widths");
  constexpr int NT = MRG ? 96 : NS * TD;               // threads per block
  constexpr int XW = PACK8 ? TX / 2 : TX;              // ring words per thread and row
  constexpr int SLOT = XW * NT;                        // ring words per input row
  constexpr int NHW = NH + (MRG ? TX : 0), NRCW = NRC + (MRG ? TX : 0); // staged per warp (the merged warp: two strips)
  constexpr int NLL = (NHW + 31) / 32, NLR = (NRCW + 31) / 32;
  constexpr int SLS = CostStage<BW, TX, MRG>::SLS, SRS = CostStage<BW, TX, MRG>::SRS;

  const int D = DT ? DT : Drt;
  const int td = threadIdx.x;
  const int lane = td & 31;
  const int wq = td >> 5;
  const bool merged = MRG && wq == 2;                               // warp-uniform
  const int strip = MRG ? (merged ? lane >> 4 : wq) : (int)threadIdx.y;
  const int d_lo = MRG ? (merged ? DCW + 2 * (lane & 15) : 2 * lane) : dbase + 2 * td;
  const int dbw = MRG ? (merged ? DCW : 0) : dbase + DCW * wq;    // first disparity of my warp
  const int xs = xblk * (NS * TX) - HW;
  const int ib = strip * TX;                                        // block index of my first hamming column
  const int ibw = merged ? 0 : ib;                                  // ... of my warp's staging buffers
  const int imax = cols - 1 - xs;
  bool live = d_lo < D;
  if (MRG) { // (the kernel's per-strip border handling, per lane)
    const int xo = (xblk * NS + strip) * TX;
    live = live && xo < cols;
    rb = xo + TX == cols;
  }
  if (dbw >= D) return;

  const int nhw = merged ? NHW : NH, nrcw = merged ? NRCW : NRC;
  int colL[NLL], colR[NLR];
#pragma unroll
  for (int k = 0; k < NLL; ++k) colL[k] = min(max(xs + ibw + lane + 32 * k, 0), cols - 1);
#pragma unroll
  for (int k = 0; k < NLR; ++k) colR[k] = min(max(xs + ibw + lane + 32 * k - DCW - 1 - dbw, 0), cols - 1);
  const uint32_t sL_s = (uint32_t)__cvta_generic_to_shared(sLb), sR_s = (uint32_t)__cvta_generic_to_shared(sRb);
  auto fetch = [&](int yin, int stage) {
    const int yc = min(max(yin, 0), rows - 1);
    const uint32_t *l = imL + (size_t)yc * cols;
    const uint32_t *r = imR + (size_t)yc * cols;
#pragma unroll
    for (int k = 0; k < NLL; ++k)
      if (lane + 32 * k < nhw)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sL_s + 4u * (uint32_t)(stage * SLS + lane + 32 * k)), "l"(l + colL[k]) : "memory");
#pragma unroll
    for (int k = 0; k < NLR; ++k)
      if (lane + 32 * k < nrcw)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sR_s + 4u * (uint32_t)(stage * SRS + lane + 32 * k)), "l"(r + colR[k]) : "memory");
  };
  auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };

  uint32_t vacc[TX]; // vacc[x] = the last emitted sum of column x (all BH rows)
#pragma unroll
  for (int x = 0; x < TX; ++x) vacc[x] = 0;
  constexpr int RS = MRG ? NT : TD; // ring stride between a thread's words (thread-private entries)
  uint32_t *myring = MRG ? ring + td : ring + (size_t)strip * XW * TD + td;
  if (BH > 1) { // rows above the band read as zero
#pragma unroll
    for (int r = 0; r < BH; ++r)
#pragma unroll
      for (int x = 0; x < XW; ++x) myring[(size_t)r * SLOT + x * RS] = 0;
  }
  const int xo0 = xs + HW + ib;
  char *prow = reinterpret_cast<char *>(outC) + (((size_t)y_begin * cols + xo0) * D + d_lo) * 2;
  const size_t rowpitch = (size_t)cols * D * 2;
  const uint32_t colpitch = (uint32_t)D * 2;

  uint32_t h[NH];                                  // h[i]: Hamming pair of column i (this row, or already the next one)
  uint32_t avn[(NH + 3) & ~3], rvn[((NH + 2) & ~1) + 2]; // staged census words; only a quad / a pair is live
  uint32_t hold = 0;
  // Hamming pairs of columns [i0, i1) of the row staged in `slot`, in ascending order
  auto ham_cols = [&](int slot, uint32_t dep, int i0, int i1) {
    const uint32_t pl = sL_s + 4u * (uint32_t)(slot * SLS + (ib - ibw));
    const uint32_t pr = sR_s + 4u * (uint32_t)(slot * SRS + (DCW - 2 * (merged ? lane & 15 : lane)) + (ib - ibw));
#pragma unroll
    for (int i = i0; i < i1; ++i) {
      // avn / rvn persist across the calls of one row (columns come in ascending order, every row
      // starts at column 0): a new left quad every 4 columns, a new right pair every 2
      if ((i & 3) == 0)
        asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(avn[i]), "=r"(avn[i + 1]), "=r"(avn[i + 2]), "=r"(avn[i + 3]) : "r"(pl + 4u * (uint32_t)i));
      if (i == 0)
        asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rvn[0]), "=r"(rvn[1]) : "r"(pr));
      if (i & 1)
        asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rvn[i + 1]), "=r"(rvn[i + 2]) : "r"(pr + 4u * (uint32_t)(i + 1)));
      uint32_t x0, x1, hv;
      asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x0) : "r"(avn[i]), "r"(rvn[i + 1]), "r"(dep));
      asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x1) : "r"(avn[i]), "r"(rvn[i]), "r"(dep));
      asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(hv) : "r"(__popc(x1)), "r"(__popc(x0)));
      if (EDGE) {
        if (i == 0) hold = 0;
        if (ib + i <= imax) hold = hv;
        h[i] = hold;
      } else {
        if (i >= NH - HW) hv = rb ? h[NH - 1 - HW] : hv;
        h[i] = hv;
      }
    }
  };
  // positions [p0, p1) of the pair stream: p < NH is column p of the current row, p >= NH column p-NH of the next
  auto produce = [&](int cslot, int nslot, uint32_t dep, int p0, int p1) {
    if (p0 < NH) ham_cols(cslot, dep, p0, p1 < NH ? p1 : NH);
    if (p1 > NH) ham_cols(nslot, dep, p0 > NH ? p0 - NH : 0, p1 - NH);
  };

  const int nin = (y_end - y_begin) + BH - 1;
#pragma unroll
  for (int s0 = 0; s0 < CSTAGES - 1; ++s0) {
    if (s0 <= nin) fetch(y_begin - HH + s0, s0);
    commit();
  }
  asm volatile("cp.async.wait_group %0;" ::"n"(CSTAGES - 2) : "memory");
  __syncwarp();
  ham_cols(0, 0u, 0, LEAD);
  int wslot = 0; // ring slot of this input row: holds row it-BH until it is overwritten
  int cslot = 0; // staging slot of input row it (row it+1 is in cslot+1)
  for (int it = 0; it < nin; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(CSTAGES - 3) : "memory");
    __syncwarp(); // row it+1 is visible to the whole warp, and everybody is done with row it-1's slot
    const int nslot = cslot + 1 == CSTAGES ? 0 : cslot + 1;
    {
      const int fb = cslot == 0 ? CSTAGES - 1 : cslot - 1; // slot of row it-1 == slot of row it+CSTAGES-1
      if (it + CSTAGES - 1 <= nin) fetch(y_begin - HH + it + CSTAGES - 1, fb);
      commit();
    }
    const bool emit = it >= BH - 1;
    uint32_t *rs = myring + (size_t)wslot * SLOT;
    uint32_t hs = 0; // horizontal sum of the window of column 0
#pragma unroll
    for (int i = 0; i < BW; ++i) hs += h[i];
    uint32_t hprev = 0, wold = 0, dep = 0;
#pragma unroll
    for (int x = 0; x < TX; ++x) {
      produce(cslot, nslot, dep, LEAD + (x * NH) / TX, LEAD + ((x + 1) * NH) / TX);
      if (BH > 1) {
        uint32_t old;
        if (!PACK8) old = rs[x * RS];
        else if (!(x & 1)) { wold = rs[(x >> 1) * RS]; old = __byte_perm(wold, 0u, 0x4140); }
        else old = __byte_perm(wold, 0u, 0x4342);
        vacc[x] = vacc[x] + hs - old;
        if (!PACK8) rs[x * RS] = hs;
        else if (x & 1) rs[(x >> 1) * RS] = __byte_perm(hprev, hs, 0x6420);
        else hprev = hs;
      } else vacc[x] = hs;
      if (live && emit && (!EDGE || xo0 + x < cols)) {
        char *dst = prow + (uint32_t)x * colpitch;
        if (!ODD_D) {
          *reinterpret_cast<uint32_t *>(dst) = vacc[x];
        } else {
          uint16_t *d16 = reinterpret_cast<uint16_t *>(dst);
          d16[0] = (uint16_t)(vacc[x] & 0xffffu);
          if (d_lo + 1 < D) d16[1] = (uint16_t)(vacc[x] >> 16);
        }
      }
      dep = hs & zmask;
      if (x + 1 < TX) hs = hs + h[x + BW] - h[x]; // window of the next column
    }
    if (emit) prow += rowpitch;
    wslot = wslot + 1 == BH ? 0 : wslot + 1;
    cslot = nslot;
  }
}


// KA: columns of extra look-ahead of the Hamming-pair stream; MINB: resident blocks per SM the register budget is sized for
template <int BW, int BH, int TX, int NS, int TD, bool PACK8, int DT, int KA, int MINB>
__global__ void __launch_bounds__(TD *NS, MINB)
cost_kernel(const uint32_t *__restrict__ cL, const uint32_t *__restrict__ cR,
            uint16_t *__restrict__ C, int rows, int cols, int D, int nchunks, int ry, uint32_t zmask) {
  constexpr int NWARP = NS * TD / 32;
  __shared__ __align__(16) uint32_t sLall[NWARP * CSTAGES * CostStage<BW, TX>::SLS];
  __shared__ __align__(16) uint32_t sRall[NWARP * CSTAGES * CostStage<BW, TX>::SRS];
  const int wib = (threadIdx.y * TD + threadIdx.x) >> 5; // warp in block
  uint32_t *sL = sLall + wib * CSTAGES * CostStage<BW, TX>::SLS;
  uint32_t *sR = sRall + wib * CSTAGES * CostStage<BW, TX>::SRS;
  extern __shared__ uint32_t ring[]; // [BH][NS][TX][TD], thread-private entries
  const int chunk = blockIdx.x % nchunks;
  const int xblk = blockIdx.x / nchunks;
  const int n = blockIdx.z;
  const int y_begin = blockIdx.y * ry;
  const int y_end = min(rows, y_begin + ry);
  const uint32_t *imL = cL + (size_t)n * rows * cols;
  const uint32_t *imR = cR + (size_t)n * rows * cols;
  uint16_t *outC = C + (size_t)n * rows * cols * D;
  // Right image border.  When the strips tile the image (cols % TX == 0) only the LAST strip is affected, and only
  // in its last BW/2 Hamming columns, which replicate the border column: a warp-uniform select on those (rb); strips
  // beyond the image have nothing to do.  Otherwise the right-most block runs the EDGE variant (a per-column hold
  // and a store bound: ~20 % more instructions, and in a one-wave grid its blocks are the tail of the kernel).
  const bool ragged = cols % TX != 0;
  const bool edge = ragged && (xblk + 1) * (NS * TX) + BW / 2 > cols;
  const int xo0 = (xblk * NS + (int)threadIdx.y) * TX; // first output column of my strip
  if (!ragged && xo0 >= cols) return;                 // warps are independent (no block barrier)
  const bool rb = !ragged && xo0 + TX == cols;
  unsigned long long *const tr = g_cost_trace;
  const int tslot = (blockIdx.y * gridDim.x + blockIdx.x) % 8192;
  if (tr && threadIdx.x == 0 && threadIdx.y == 0) {
    unsigned sm;
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    tr[4 * tslot] = cost_gtimer(); tr[4 * tslot + 2] = sm;
  }
  if (DT == 0 && (D & 1)) // odd D (generic aggregation path only): 16-bit stores, keep one slow variant
    cost_band_s<BW, BH, TX, NS, TD, true, PACK8, true, 0, KA>(imL, imR, outC, rows, cols, D, chunk * 2 * TD, xblk, y_begin, y_end, ring, sL, sR, zmask, rb);
  else if (edge)
    cost_band_s<BW, BH, TX, NS, TD, true, PACK8, false, DT, KA>(imL, imR, outC, rows, cols, D, chunk * 2 * TD, xblk, y_begin, y_end, ring, sL, sR, zmask, rb);
  else
    cost_band_s<BW, BH, TX, NS, TD, false, PACK8, false, DT, KA>(imL, imR, outC, rows, cols, D, chunk * 2 * TD, xblk, y_begin, y_end, ring, sL, sR, zmask, rb);
  if (tr && threadIdx.x == 0 && threadIdx.y == 0) tr[4 * tslot + 1] = cost_gtimer();
}

// D = 96 with strips that tile the image: three warps per block for two strips (cost_band_s, MRG)
template <int BW, int BH, int TX, bool PACK8, int KA, int MINB>
__global__ void __launch_bounds__(96, MINB)
cost96_kernel(const uint32_t *__restrict__ cL, const uint32_t *__restrict__ cR,
              uint16_t *__restrict__ C, int rows, int cols, int ry, uint32_t zmask) {
  using Stage = CostStage<BW, TX, true>;
  __shared__ __align__(16) uint32_t sLall[3 * CSTAGES * Stage::SLS];
  __shared__ __align__(16) uint32_t sRall[3 * CSTAGES * Stage::SRS];
  const int wib = threadIdx.x >> 5;
  extern __shared__ uint32_t ring[]; // [BH][TX (/2)][96], thread-private entries
  const int n = blockIdx.z;
  const int y_begin = blockIdx.y * ry;
  const int y_end = min(rows, y_begin + ry);
  cost_band_s<BW, BH, TX, 2, 64, false, PACK8, false, 96, KA, true>(
      cL + (size_t)n * rows * cols, cR + (size_t)n * rows * cols, C + (size_t)n * rows * cols * 96, rows, cols, 96, 0,
      (int)blockIdx.x, y_begin, y_end, ring, sLall + wib * CSTAGES * Stage::SLS, sRall + wib * CSTAGES * Stage::SRS, zmask, false);
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

// Rows per band.  As many bands as fit in ONE resident wave (a second, partial wave would double the kernel
// time), at least 12 rows each so that the BH-1 warm-up rows stay a bounded overhead.  Grids that exceed one wave
// anyway -- or would leave more than a fifth of it empty with whole bands (e.g. 16 environments of 848x480: 432
// blocks for 740 slots) -- use ~64-row bands (~128-row bands for large batches) and several waves.
static int cost_band_rows(long capacity, long columns, int rows) {
  long bands = capacity / columns;
  const long max_bands = (rows + 11) / 12; // (ROI 640x360: 12-row bands 33 us, 24-row bands 37.5 us, 8-row bands 30 us; frame rate at three lanes unchanged)
  if (bands >= max_bands) bands = max_bands;
  else if (bands < 1 || 5 * bands * columns < 4 * capacity) {
    bands = (rows + 127) / 128;                                            // (6 warm-up rows per band: 128-row bands when
    if (bands * columns < 4 * capacity) bands = (rows + 63) / 64;          //  that still makes four waves or more)
  }
  return (int)((rows + bands - 1) / bands);
}

template <int BW, int BH, int TX, int NS, int TD, bool PACK8, int DT = 0, int KA = 4, int MINB = 1>
static cudaError_t launch_cfg(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                              int cols, int D, cudaStream_t st) {
  auto k = cost_kernel<BW, BH, TX, NS, TD, PACK8, DT, KA, MINB>;
  const size_t smem = (size_t)BH * NS * TX * TD * sizeof(uint32_t) / (PACK8 ? 2 : 1);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, TD * NS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int nchunks = (D + 2 * TD - 1) / (2 * TD);
  const long xb = (long)((cols + NS * TX - 1) / (NS * TX)) * nchunks;
  const long capacity = (long)sm_count() * per_sm;
  const int ry = cost_band_rows(capacity, xb * N, rows);
  dim3 grid((unsigned)xb, (unsigned)((rows + ry - 1) / ry), (unsigned)N);
  k<<<grid, dim3(TD, NS), smem, st>>>(cL, cR, C, rows, cols, D, nchunks, ry, 0u); // 0u: the opaque zero of the software pipeline
  return cudaGetLastError();
}

template <int BW, int BH, int TX, bool PACK8, int KA, int MINB>
static cudaError_t launch_cfg96(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows, int cols, cudaStream_t st) {
  auto k = cost96_kernel<BW, BH, TX, PACK8, KA, MINB>;
  const size_t smem = (size_t)BH * TX * 96 * sizeof(uint32_t) / (PACK8 ? 2 : 1);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 96, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long xb = (cols + 2 * TX - 1) / (2 * TX);
  const long capacity = (long)sm_count() * per_sm;
  const int ry = cost_band_rows(capacity, xb * N, rows);
  dim3 grid((unsigned)xb, (unsigned)((rows + ry - 1) / ry), (unsigned)N);
  k<<<grid, 96, smem, st>>>(cL, cR, C, rows, cols, ry, 0u);
  return cudaGetLastError();
}

template <int BW, int BH>
static cudaError_t launch_fast(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                               int cols, int D, int bits, cudaStream_t st) {
  constexpr int TX = 16;
  // a BW-wide Hamming sum fits one byte: half-size ring (BH == 1 reads back the word it is writing)
  const bool pack8 = BH > 1 && BW * bits <= 255;
  if constexpr (BW == 7 && BH == 7) if (pack8) { // the stock block size: compile-time D for the usual disparity ranges
    // 32 columns per thread (38/32 instead of 22/16 Hamming columns per output column) where the strips tile the usual
    // image widths; D = 96 is the 848-column sensor (848 % 32 = 16 would put every right-most block on the EDGE path).
    // 159 registers -> three blocks (12 warps) per SM; D = 96: 118 registers, four blocks.  (Two blocks per SM with the
    // same schedule: C1 stage 62 -> 66 us; frame rate with three lanes and the C5 sweep with two pipelines unchanged.)
    if (D == 64) return launch_cfg<BW, BH, 32, 4, 32, true, 64, 4, 3>(cL, cR, C, N, rows, cols, D, st);
    // (three warps for two strips: C3, 64 envs, cost 1.52 -> 1.24 ms; ragged widths keep four warps, two of them half empty)
    if (D == 96 && cols % TX == 0) return launch_cfg96<BW, BH, TX, true, 4, 5>(cL, cR, C, N, rows, cols, st);
    if (D == 96) return launch_cfg<BW, BH, TX, 2, 64, true, 96, 4, 4>(cL, cR, C, N, rows, cols, D, st);
    if (D == 128) return launch_cfg<BW, BH, 32, 2, 64, true, 128, 4, 3>(cL, cR, C, N, rows, cols, D, st);
    if (D == 256) return launch_cfg<BW, BH, 32, 2, 64, true, 256, 4, 3>(cL, cR, C, N, rows, cols, D, st);
  }
  if (D <= 64)
    return pack8 ? launch_cfg<BW, BH, TX, 4, 32, true>(cL, cR, C, N, rows, cols, D, st)
                 : launch_cfg<BW, BH, TX, 4, 32, false>(cL, cR, C, N, rows, cols, D, st);
  return pack8 ? launch_cfg<BW, BH, TX, 2, 64, true>(cL, cR, C, N, rows, cols, D, st)
               : launch_cfg<BW, BH, TX, 2, 64, false>(cL, cR, C, N, rows, cols, D, st);
}

} // namespace ssb
extern "C" int ssb_debug_set_cost_trace(void *device_buffer) {
  return (int)cudaMemcpyToSymbol(ssb::g_cost_trace, &device_buffer, sizeof(device_buffer));
}
namespace ssb {

cudaError_t launch_cost(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                        int cols, int D, int bw, int bh, int bits, cudaStream_t st) {
  if (N > 65535) return cudaErrorInvalidValue;
  if (bw == 7 && bh == 7) return launch_fast<7, 7>(cL, cR, C, N, rows, cols, D, bits, st);
  if (bw == 1 && bh == 1) return launch_fast<1, 1>(cL, cR, C, N, rows, cols, D, bits, st);
  if (bw == 3 && bh == 3) return launch_fast<3, 3>(cL, cR, C, N, rows, cols, D, bits, st);
  if (bw == 5 && bh == 5) return launch_fast<5, 5>(cL, cR, C, N, rows, cols, D, bits, st);
  if (bw == 9 && bh == 9) return launch_fast<9, 9>(cL, cR, C, N, rows, cols, D, bits, st);
  const size_t total = (size_t)N * rows * cols * D;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cost_generic_kernel<<<blocks, 256, 0, st>>>(cL, cR, C, rows, cols, D, bw, bh, total);
  return cudaGetLastError();
}

} // namespace ssb
