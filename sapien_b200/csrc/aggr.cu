// aggr.cu -- SGM path aggregation (4 paths), blend and winner-takes-all for sm_100a.
//
// Replaces the reference kernels aggrLeft2Right / aggrRight2Left / aggrTop2Bottom /
// aggrBottom2Top (3rd_party/simsense/src/aggr.cu:29-230) and winnerTakesAll
// (src/wta.cu:170-214).  Semantics reproduced exactly; the
// organisation is new:
//
//  * one WARP walks one path (an image row or column); a lane owns DPL = 2*NR consecutive
//    disparities packed as u16x2 registers.  A path step is ~25 instructions and needs no
//    __syncthreads and no atomics (the reference: one thread per disparity, 3 barriers and one
//    shared atomicMin per step);
//  * d-1 / d+1 neighbours: funnel shifts inside the lane, one shuffle each way across lanes, and a
//    per-lane byte-permute selector that folds the d=0 / d=D-1 border rule into the same PRMT;
//    the running minimum is vmin + half-swap + ONE redux.sync (CREDUX) that already returns
//    m|m<<16; the min/add chain is Blackwell DPX (VIADDMNMX.U16x2, VIMNMX.U16x2, VIMNMX3.U16x2);
//  * the cost volume (and the partial sums of earlier passes) are streamed into shared memory by the
//    copy engines, completion on an mbarrier, in chunks of K path steps, NCH chunks deep per warp:
//    horizontal paths take one 4 KB bulk copy per chunk (cp.async.bulk / UBLKCP); vertical paths take
//    ONE 2-D tensor copy per stream and chunk (cp.async.bulk.tensor / UTMALDG, box = K rows x one
//    pixel's D costs, a CUtensorMap per volume in the kernel arguments) when a pixel's costs are whole
//    128-byte lines, else K bulk row pieces issued by K lanes.  No per-step address arithmetic, no
//    LDGSTS issue cost; a warp keeps (NCH-1)*K*D*2 bytes per stream in flight;
//  * pass order is  right->left || top->bottom (two streams)  ->  bottom->top (+L1+L2)  ->  left->right.
//    The last pass forms LAll(y,x,:) = (L0+L1+L2+L3)/4 in registers.  Per step it only records the
//    packed (min,argmin) key and advances the right-disparity recurrence along the diagonal,
//    T_x(d) = min(T_{x-1}(d-1), LAll(x,d)); the LAll row goes to a 32-pixel shared tile and every 32
//    steps the warp switches to ONE LANE PER PIXEL for uniqueness + sub-pixel (packed 3-input
//    min scan of the row with the d*-1..d*+1 window masked), so the winner-takes-all costs ~4
//    instructions per pixel instead of a block per pixel.  LAll is never written to HBM.
//
// HBM traffic: 2V + 2V + 4V + 2V = 10 V for aggregation + WTA, against ~13 V in the reference
// (V = one u16 volume).
#include "common.cuh"
#include "kernels.h"

#include <cuda.h> // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

namespace ssb {

struct AggrArgs {
  const uint16_t *C;
  const uint16_t *aux0, *aux1;
  uint16_t *out;  // MODE 0: L ; MODE 1: L+aux0+aux1
  uint16_t *dbg0; // MODE 1: L3 ; MODE 2: L0      (keep_stages only)
  uint16_t *dbg1; // MODE 2: LAll                 (keep_stages only)
  float *dispL;
  uint16_t *dispR;
  int N, rows, cols, D;
  int vertical, reverse;
  uint32_t P1P1, P2P2;
  int uniq;
  int nsm; // SM count
  // 12-bit storage of the two plain path volumes (out of a MODE 0 pass, aux0 / aux1 of the MODE 1 pass): a pixel's D
  // costs take 1.5 D bytes -- D low bytes, then D/2 bytes of high nibbles -- instead of 2 D.  Storage only: the
  // arithmetic stays on u16x2 registers, so results are bit-identical by construction.  Valid while every path cost is
  // < 4096, i.e. cmax + P2 < 4096 (aggr_pack12_supported).
  int pack12;
  // final pass only: after the columns < seg_end[k] of a row are finished (disparities written),
  // its consumer warp bumps progress[k]; a stream can wait on the counter (cuStreamWaitValue32) and
  // post-process / copy those columns while the pass is still running
  uint32_t *progress;
  int nseg, seg_end[6];
  int tma;  // vertical passes: rows of a chunk arrive as ONE 2-D tensor copy per stream (tm[] valid)
  alignas(64) CUtensorMap tm[3]; // C, aux0, aux1 viewed as [N*rows][cols*D] u16, box = [K][D]
};

// ---- optional per-block timeline (debug/profiling aid, tools/trace_aggr.py): when a buffer is
// registered, lane 0 of every path warp records {start ns, end ns, SM id, start clock} --------------
__device__ unsigned long long *g_trace = nullptr;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %smid;" : "=r"(r));
  return r;
}
constexpr int TRACE_STRIDE = 8192; // records per kernel slot

// ---- mbarrier + bulk-copy (TMA 1-D) primitives -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_g2s_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int NR> __device__ __forceinline__ void lds_vec(const void *p, uint32_t (&r)[NR]) {
  if constexpr (NR == 1) {
    r[0] = *reinterpret_cast<const uint32_t *>(p);
  } else if constexpr (NR == 2) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    r[0] = v.x; r[1] = v.y;
  } else if constexpr (NR == 3) { // 12-byte lanes (D = 96 on a half-warp): three words, stride 3 -> no bank conflict
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
    r[0] = q[0]; r[1] = q[1]; r[2] = q[2];
  } else {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i) {
      const uint4 v = reinterpret_cast<const uint4 *>(p)[i];
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  }
}
template <int NR> __device__ __forceinline__ void st_vec(void *p, const uint32_t (&r)[NR]) {
  if constexpr (NR == 1) {
    *reinterpret_cast<uint32_t *>(p) = r[0];
  } else if constexpr (NR == 2) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(r[0], r[1]);
  } else if constexpr (NR == 3) {
    uint32_t *q = reinterpret_cast<uint32_t *>(p);
    q[0] = r[0]; q[1] = r[1]; q[2] = r[2];
  } else {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  }
}


// ---- 12-bit packed pixels (AggrArgs::pack12) ----------------------------------------------------------
// A lane owns 2 NR consecutive disparities: 2 NR low bytes at piece + 2 NR lane and NR bytes of high nibbles at
// piece + D + NR lane.  Two registers (4 values) at a time: low bytes b3 b2 b1 b0, nibbles n3 n2 n1 n0.
__device__ __forceinline__ void unpack12_quad(uint32_t lo, uint32_t hi16, uint32_t &r0, uint32_t &r1) {
  uint32_t x = (hi16 | (hi16 << 8)) & 0x00ff00ffu; // 00 n3n2 00 n1n0
  x = (x | (x << 4)) & 0x0f0f0f0fu;                 // 0n3 0n2 0n1 0n0
  r0 = __byte_perm(lo, x, 0x5140);                  // (b0 | n0 << 8) | (b1 | n1 << 8) << 16
  r1 = __byte_perm(lo, x, 0x7362);
}
__device__ __forceinline__ void pack12_quad(uint32_t r0, uint32_t r1, uint32_t &lo, uint32_t &hi16) {
  lo = __byte_perm(r0, r1, 0x6420);
  uint32_t x = __byte_perm(r0, r1, 0x7531);         // 0n3 0n2 0n1 0n0
  x = (x | (x >> 4)) & 0x00ff00ffu;
  hi16 = (x | (x >> 8)) & 0xffffu;
}
template <int NR> __device__ __forceinline__ void lds_unpack12(const unsigned char *piece, int D, int lane, uint32_t (&r)[NR]) {
  if constexpr (NR == 1) {
    const uint32_t lo = *reinterpret_cast<const uint16_t *>(piece + 2 * lane);
    const uint32_t hi = piece[D + lane];
    const uint32_t x = (hi | (hi << 4)) & 0x0f0fu;
    r[0] = __byte_perm(lo, x, 0x5140);
  } else {
    uint32_t lo[NR / 2], hi[(NR + 3) / 4];
    const unsigned char *pl = piece + 2 * NR * lane, *ph = piece + D + NR * lane;
    if constexpr (NR == 2) { lo[0] = *reinterpret_cast<const uint32_t *>(pl); hi[0] = *reinterpret_cast<const uint16_t *>(ph); }
    else if constexpr (NR == 4) { const uint2 v = *reinterpret_cast<const uint2 *>(pl); lo[0] = v.x; lo[1] = v.y; hi[0] = *reinterpret_cast<const uint32_t *>(ph); }
    else {
#pragma unroll
      for (int i = 0; i < NR / 8; ++i) {
        const uint4 v = reinterpret_cast<const uint4 *>(pl)[i];
        lo[4 * i] = v.x; lo[4 * i + 1] = v.y; lo[4 * i + 2] = v.z; lo[4 * i + 3] = v.w;
        const uint2 h = reinterpret_cast<const uint2 *>(ph)[i];
        hi[2 * i] = h.x; hi[2 * i + 1] = h.y;
      }
    }
#pragma unroll
    for (int q = 0; q < NR / 2; ++q) unpack12_quad(lo[q], (hi[q / 2] >> (16 * (q & 1))) & 0xffffu, r[2 * q], r[2 * q + 1]);
  }
}
template <int NR> __device__ __forceinline__ void st_pack12(unsigned char *piece, int D, int lane, const uint32_t (&r)[NR]) {
  if constexpr (NR == 1) {
    *reinterpret_cast<uint16_t *>(piece + 2 * lane) = (uint16_t)__byte_perm(r[0], 0u, 0x4420);
    const uint32_t x = __byte_perm(r[0], 0u, 0x4431); // 0n1 0n0
    piece[D + lane] = (unsigned char)((x | (x >> 4)) & 0xffu);
  } else {
    uint32_t lo[NR / 2], hi[(NR + 3) / 4];
#pragma unroll
    for (int i = 0; i < (NR + 3) / 4; ++i) hi[i] = 0;
#pragma unroll
    for (int q = 0; q < NR / 2; ++q) {
      uint32_t h;
      pack12_quad(r[2 * q], r[2 * q + 1], lo[q], h);
      hi[q / 2] |= h << (16 * (q & 1));
    }
    unsigned char *pl = piece + 2 * NR * lane, *ph = piece + D + NR * lane;
    if constexpr (NR == 2) { *reinterpret_cast<uint32_t *>(pl) = lo[0]; *reinterpret_cast<uint16_t *>(ph) = (uint16_t)hi[0]; }
    else if constexpr (NR == 4) { *reinterpret_cast<uint2 *>(pl) = make_uint2(lo[0], lo[1]); *reinterpret_cast<uint32_t *>(ph) = hi[0]; }
    else {
#pragma unroll
      for (int i = 0; i < NR / 8; ++i) {
        reinterpret_cast<uint4 *>(pl)[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
        reinterpret_cast<uint2 *>(ph)[i] = make_uint2(hi[2 * i], hi[2 * i + 1]);
      }
    }
  }
}

// One SGM path step on packed u16x2 registers (aggr.cu:39-76 of the reference):
//   L'(d) = C(d) + min(L(d), L(d-1)+P1, L(d+1)+P1, m+P2) - m,   m = min_k L(k).
// selUp / selDn are per-lane PRMT selectors: 0x5432 = take the neighbour lane's half, 0x5454 /
// 0x3232 (first / last disparity) = repeat the own value, which makes the missing neighbour
// harmless (L(d)+P1 never beats L(d)).  L == 0 everywhere reproduces the first-pixel rule L = C.
// Minimum over the 16 lanes of the caller's half-warp, for both halves at once (two row paths share a warp, see
// aggr_wta2_kernel): CREDUX reduces the whole warp into a uniform register, so each half takes its turn with the
// other half contributing the neutral element.
__device__ __forceinline__ uint32_t halfwarp_min(uint32_t t, bool upper) {
  const uint32_t lo = __reduce_min_sync(FULL, upper ? 0xffffffffu : t);
  const uint32_t hi = __reduce_min_sync(FULL, upper ? t : 0xffffffffu);
  return upper ? hi : lo;
}

template <int NR, bool HALF = false>
__device__ __forceinline__ void sgm_step(uint32_t (&L)[NR], const uint32_t (&c)[NR], uint32_t P1P1,
                                         uint32_t P2P2, uint32_t selUp, uint32_t selDn, bool upper = false) {
  uint32_t t = L[0];
#pragma unroll
  for (int j = 1; j < NR; ++j) t = __vminu2(t, L[j]);
  t = __vminu2(t, __byte_perm(t, t, 0x1032));        // both halves = lane minimum
  const uint32_t mm = HALF ? halfwarp_min(t, upper) : __reduce_min_sync(FULL, t); // m | m << 16
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
    nl[j] = v - mm + c[j]; // per half: v >= m, and the sum stays below 65536 in the fast regime
  }
#pragma unroll
  for (int j = 0; j < NR; ++j) L[j] = nl[j];
}

// MODE 0: plain path (out = L).  MODE 1: out = L + aux0 + aux1.  MODE 2: left->right path,
// LAll = (L + aux0)/4, winner-takes-all (horizontal forward only; producer + consumer warp).
template <int MODE> __host__ __device__ constexpr int nstream() { return MODE == 0 ? 1 : (MODE == 1 ? 3 : 2); }

struct AggrSmem { // per-path layout (bytes); PIECE = D*2
  int ring, tile, gk, rb, total;
};
constexpr int WTA_TILES = 2; // LAll tiles (32 pixels each) between the SGM warp and the WTA warp
template <int MODE, int K, int NCH> __host__ __device__ inline AggrSmem aggr_smem(int D) {
  AggrSmem s;
  const int piece = D * 2;
  s.ring = 128; // (tensor copies want a 128-byte aligned destination) mbarriers first: NCH bulk-copy barriers, then (MODE 2) WTA_TILES full + WTA_TILES empty
  s.tile = s.ring + nstream<MODE>() * NCH * K * piece;
  s.gk = s.tile + (MODE == 2 ? (WTA_TILES * 32 + 1) * (piece + 16) : 0); // +1 row: the consumer prefetches one row ahead
  s.rb = s.gk + (MODE == 2 ? 32 * 4 : 0);
  s.total = s.rb + (MODE == 2 ? 32 * 2 : 0);
  s.total = (s.total + 127) & ~127;
  return s;
}

// The bulk-copy ring of one path: NS streams x NCH chunk slots x K steps x PIECE bytes.
template <int NS, int K, int NCH> struct PathRing {
  uint32_t bar0, ring_s;
  const char *g[3];
  long sstride;
  int steps, PIECE, STREAM, lane;
  // streams 1.. (aux0, aux1) when they are 12-bit packed: PX = 1.5 D bytes per pixel (rows of a chunk lie at pitch PX
  // inside the chunk slot), xstride = byte stride between steps, c0x = tensor column of the path; else the stream-0 values
  int PX;
  long xstride;
  int c0x;
  bool vertical, hrev;
  // 2-D tensor-copy mode of vertical paths: box = K rows x one pixel's D costs
  const CUtensorMap *tm;
  int tma, c0, row0, vrev; // element column of the path, first row (n*rows [+ rows-1 when reversed]), direction
  __device__ __forceinline__ void init() const {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) mbar_init(bar0 + 8 * i, 1);
    }
  }
  // all lanes call; chunk ci -> slot ci % NCH
  __device__ __forceinline__ void issue(int ci) const {
    const int slot = ci % NCH;
    const int s0 = ci * K;
    const int kc = min(K, steps - s0);
    const uint32_t bar = bar0 + 8 * slot;
    const uint32_t dst = ring_s + (uint32_t)(slot * K * PIECE);
    if (tma) {
      // One instruction per stream instead of K row pieces (each UBLKCP of a divergent lane costs an
      // ELECT / R2UR / branch round: ~10 instructions per piece, 3 pieces per step in the 3-stream
      // pass).  The box is always K rows: rows past the path's end belong to the neighbouring
      // environment or are out of bounds (zero filled) and are never consumed.
      if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(K * (PIECE + (NS - 1) * PX)));
        const int c1 = vrev ? row0 - s0 - (K - 1) : row0 + s0;
#pragma unroll
        for (int i = 0; i < NS; ++i) tma_g2s_2d(dst + i * STREAM, tm + i, i ? c0x : c0, c1, bar);
      }
      return;
    }
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(kc * (PIECE + (NS - 1) * PX)));
    __syncwarp();
    if (vertical) {
      if (lane < kc) {
        bulk_g2s(dst + (uint32_t)(lane * PIECE), g[0] + (long)(s0 + lane) * sstride, (uint32_t)PIECE, bar);
#pragma unroll
        for (int i = 1; i < NS; ++i)
          bulk_g2s(dst + i * STREAM + (uint32_t)(lane * PX), g[i] + (long)(s0 + lane) * xstride, (uint32_t)PX, bar);
      }
    } else if (lane == 0) {
      const long at = hrev ? s0 + kc - 1 : s0;
      bulk_g2s(dst, g[0] + at * sstride, (uint32_t)(kc * PIECE), bar);
#pragma unroll
      for (int i = 1; i < NS; ++i) bulk_g2s(dst + i * STREAM, g[i] + at * xstride, (uint32_t)(kc * PX), bar);
    }
  }
};

struct PathGeom {
  int n, q, steps;
  size_t e0;    // element offset of path step 0, disparity 0
  long sstride; // signed byte stride between steps
};
__device__ __forceinline__ PathGeom path_geom(const AggrArgs &a, long path) {
  PathGeom g;
  const int per_env = a.vertical ? a.cols : a.rows;
  g.n = (int)(path / per_env);
  g.q = (int)(path - (long)g.n * per_env);
  g.steps = a.vertical ? a.rows : a.cols;
  const size_t env = (size_t)g.n * a.rows * a.cols * a.D;
  g.e0 = env + (a.vertical ? (size_t)g.q * a.D : (size_t)g.q * a.cols * a.D);
  g.sstride = a.vertical ? (long)a.cols * a.D * 2 : (long)a.D * 2;
  if (a.reverse) { g.e0 += (size_t)(g.steps - 1) * (size_t)(g.sstride / 2); g.sstride = -g.sstride; }
  return g;
}

// ---- MODE 0 / 1: one independent warp per path ---------------------------------------------------
// PACK: the plain path volumes are stored 12-bit packed (AggrArgs::pack12): MODE 0 packs its output, MODE 1 unpacks
// aux0 / aux1; the cost volume and the MODE 1 output (a sum of three paths, up to 14 bits) stay u16.
template <int NR, int MODE, bool PARTIAL, bool DBG, int K, int NCH, bool PACK>
__global__ void __launch_bounds__(128) aggr_kernel(const __grid_constant__ AggrArgs a) {
  static_assert(MODE == 0 || MODE == 1, "plain passes only");
  static_assert(!(PACK && DBG), "stage materialisation keeps u16 volumes");
  constexpr int DPL = 2 * NR;
  constexpr int NS = nstream<MODE>();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const long path = (long)blockIdx.x * wpb + warp;
  if (path >= (long)a.N * (a.vertical ? a.cols : a.rows)) return; // warps are independent: no block barrier
  const PathGeom pg = path_geom(a, path);
  const int steps = pg.steps;
  const int D = PARTIAL ? a.D : 64 * NR; // whole lanes: D is a compile-time constant, ring offsets become immediates
  const int PIECE = D * 2;
  const AggrSmem lay = aggr_smem<MODE, K, NCH>(D);
  unsigned char *wsm = smem_raw + (size_t)warp * lay.total;
  PathRing<NS, K, NCH> pr;
  pr.bar0 = (uint32_t)__cvta_generic_to_shared(wsm);
  pr.ring_s = pr.bar0 + lay.ring;
  pr.g[0] = reinterpret_cast<const char *>(a.C + pg.e0);
  constexpr bool PACKIN = PACK && MODE == 1; // aux streams arrive packed
  const int PX = PACKIN ? (D * 3) / 2 : PIECE;
  const size_t xoff = PACKIN ? pg.e0 * 3 / 2 : pg.e0 * 2; // byte offset of the path's first pixel in an aux volume
  pr.g[1] = NS > 1 ? reinterpret_cast<const char *>(a.aux0) + xoff : nullptr;
  pr.g[2] = NS > 2 ? reinterpret_cast<const char *>(a.aux1) + xoff : nullptr;
  pr.sstride = pg.sstride; pr.steps = steps; pr.PIECE = PIECE; pr.STREAM = NCH * K * PIECE; pr.lane = lane;
  pr.PX = PX; pr.xstride = PACKIN ? pg.sstride * 3 / 4 : pg.sstride;
  pr.vertical = a.vertical != 0; pr.hrev = !a.vertical && a.reverse;
  pr.tm = a.tm; pr.tma = a.tma && a.vertical; pr.vrev = a.reverse;
  pr.c0 = pg.q * D; pr.row0 = pg.n * a.rows + (a.reverse ? a.rows - 1 : 0);
  pr.c0x = PACKIN ? pg.q * ((D * 3) / 4) : pr.c0;
  const unsigned char *ring = wsm + lay.ring;
  const int STREAM = pr.STREAM;
  const bool hrev = pr.hrev || (pr.tma && a.reverse); // chunk rows lie in ascending address order: walk them backwards

  const int nact = D / DPL; // active lanes
  const bool active = !PARTIAL || lane < nact;
  const int loff = lane * NR * 4; // byte offset of this lane inside a piece
  const uint32_t selUp = lane == 0 ? 0x5454u : 0x5432u;
  const uint32_t selDn = lane == nact - 1 ? 0x3232u : 0x5432u;

  pr.init();
  if (lane == 0) {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  unsigned long long *const tr = g_trace;
  const int tslot = (MODE == 1 ? 2 : a.vertical) * TRACE_STRIDE + (int)(path % TRACE_STRIDE);
  if (tr && lane == 0) { tr[4 * tslot] = gtimer(); tr[4 * tslot + 2] = smid(); }
  const int nchunks = (steps + K - 1) / K;
  for (int ci = 0; ci < NCH - 1 && ci < nchunks; ++ci) pr.issue(ci);

  uint32_t L[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) L[r] = active ? 0u : 0xffffffffu;
  constexpr bool PACKOUT = PACK && MODE == 0; // the output volume is written packed
  char *pOut = PACKOUT ? reinterpret_cast<char *>(a.out) + pg.e0 * 3 / 2 : reinterpret_cast<char *>(a.out + pg.e0) + loff;
  char *pDbg0 = DBG ? reinterpret_cast<char *>(a.dbg0 + pg.e0) + loff : nullptr;
  const long sstride = pg.sstride;
  const long ostride = PACKOUT ? pg.sstride * 3 / 4 : pg.sstride;

  // pc: this lane's slice of the cost row; px: start of the same step's row in the first aux stream (PACKIN only)
  auto step = [&](const unsigned char *pc, const unsigned char *px) {
    uint32_t c[NR], x0[NR], x1[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { c[r] = 0; x0[r] = 0; x1[r] = 0; }
    if (active) {
      lds_vec<NR>(pc, c);
      if (PACKIN) {
        if (NS > 1) lds_unpack12<NR>(px, D, lane, x0);
        if (NS > 2) lds_unpack12<NR>(px + STREAM, D, lane, x1);
      } else {
        if (NS > 1) lds_vec<NR>(pc + STREAM, x0);
        if (NS > 2) lds_vec<NR>(pc + 2 * STREAM, x1);
      }
    }
    sgm_step<NR>(L, c, a.P1P1, a.P2P2, selUp, selDn);
    if (PARTIAL) {
#pragma unroll
      for (int r = 0; r < NR; ++r) L[r] = active ? L[r] : 0xffffffffu;
    }
    if (active) {
      if (MODE == 0) {
        if (PACKOUT) st_pack12<NR>(reinterpret_cast<unsigned char *>(pOut), D, lane, L);
        else st_vec<NR>(pOut, L);
      } else {
        uint32_t o[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) o[r] = L[r] + x0[r] + x1[r];
        st_vec<NR>(pOut, o);
        if (DBG) st_vec<NR>(pDbg0, L);
      }
    }
    pOut += ostride;
    if (DBG) pDbg0 += sstride;
  };

  int slot = 0;
  uint32_t parity = 0;
  for (int ci = 0; ci < nchunks; ++ci) {
    mbar_wait(pr.bar0 + 8 * slot, parity);
    if (ci + NCH - 1 < nchunks) pr.issue(ci + NCH - 1); // refills the slot consumed one chunk ago
    const int kc = min(K, steps - ci * K);
    const unsigned char *pc = ring + slot * K * PIECE + loff;
    const unsigned char *px = ring + STREAM + slot * K * PIECE; // (packed aux rows lie at pitch PX inside the slot)
    if (kc == K) {
      if (hrev) {
        pc += (K - 1) * PIECE; px += (K - 1) * PX;
#pragma unroll
        for (int k = 0; k < K; ++k) { step(pc, px); pc -= PIECE; px -= PX; }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) { step(pc, px); pc += PIECE; px += PX; }
      }
    } else {
      const int dp = hrev ? -PIECE : PIECE, dx = hrev ? -PX : PX;
      if (hrev) { pc += ((pr.tma ? K : kc) - 1) * PIECE; px += ((pr.tma ? K : kc) - 1) * PX; } // a tensor box is always K rows, a bulk piece kc steps
      for (int k = 0; k < kc; ++k) { step(pc, px); pc += dp; px += dx; }
    }
    if (++slot == NCH) { slot = 0; parity ^= 1; }
  }
  if (tr && lane == 0) tr[4 * tslot + 1] = gtimer();
}

// Uniqueness test + sub-pixel interpolation of ONE pixel by one lane (wta.cu:164-168,203): `row` = the pixel's D
// blended costs in shared memory (the d*-1..d*+1 window is overwritten), gk = min << 16 | argmin.
__device__ __forceinline__ float wta_pixel(uint16_t *row, uint32_t gk, int D, int k100u) {
  const int m = (int)(gk >> 16), ds = (int)(gk & 0xffffu);
  int y0 = 0, y2 = 0;
  if (ds > 0) y0 = row[ds - 1];
  if (ds < D - 1) y2 = row[ds + 1];
  bool ok;
  if (k100u > 0) {
    // unique <=> min over d outside [d*-1, d*+1] of LAll(d)*(100-u) >= m*100 (wta.cu:203)
    if (ds > 0) row[ds - 1] = 0xffffu;
    row[ds] = 0xffffu;
    if (ds < D - 1) row[ds + 1] = 0xffffu;
    uint32_t mn0 = 0xffffffffu, mn1 = 0xffffffffu, mn2 = 0xffffffffu, mn3 = 0xffffffffu;
    const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
    int i = 0;
    for (; i + 4 <= D / 8; i += 4) { // 4 independent chains: the loads of one round overlap
      const uint4 v0 = r4[i], v1 = r4[i + 1], v2 = r4[i + 2], v3 = r4[i + 3];
      mn0 = __vimin3_u16x2(mn0, v0.x, v0.y); mn1 = __vimin3_u16x2(mn1, v1.x, v1.y);
      mn2 = __vimin3_u16x2(mn2, v2.x, v2.y); mn3 = __vimin3_u16x2(mn3, v3.x, v3.y);
      mn0 = __vimin3_u16x2(mn0, v0.z, v0.w); mn1 = __vimin3_u16x2(mn1, v1.z, v1.w);
      mn2 = __vimin3_u16x2(mn2, v2.z, v2.w); mn3 = __vimin3_u16x2(mn3, v3.z, v3.w);
    }
    for (; i < D / 8; ++i) {
      const uint4 v = r4[i];
      mn0 = __vimin3_u16x2(mn0, v.x, v.y);
      mn1 = __vimin3_u16x2(mn1, v.z, v.w);
    }
    const uint32_t mn = __vimin3_u16x2(mn0, mn1, __vminu2(mn2, mn3));
    const int m2 = (int)min(mn & 0xffffu, mn >> 16);
    ok = m2 * k100u >= m * 100;
  } else { // uniqueness_ratio >= 100: the product test is not monotone; evaluate it literally
    ok = true;
    for (int d = 0; d < D; ++d) {
      const int dd = d - ds;
      ok = ok && ((int)row[d] * k100u >= m * 100 || (dd >= -1 && dd <= 1));
    }
  }
  float disp = (float)ds;
  if (!ok) {
    disp = -1.0f;
  } else if (ds != 0 && ds != D - 1) {
    // wta.cu:164-168 computes (double)(y2-y0) / (2.0*(double)(y0-2*y1+y2)) and rounds to float.
    // Numerator and denominator are integers < 2^20, so ONE correctly rounded float division
    // gives the same bits (no double-rounding case exists for |num|,den < 2^24).
    disp = (float)ds - __fdiv_rn((float)(y2 - y0), (float)(2 * (y0 - 2 * m + y2)));
  }
  return disp;
}

// ---- MODE 2: left->right path + blend + winner-takes-all, two warps per image row ---------------
// Warp 0 (producer) walks the path: SGM step, LAll = (L0 + S3)/4, row -> shared tile.  Warp 1
// (consumer) follows one tile (32 pixels) behind: packed (min,argmin) keys, the right-disparity
// diagonal recurrence, and after each tile the one-lane-per-pixel uniqueness / sub-pixel phase.
// Hand-over through two full/empty mbarrier pairs, so the serial SGM chain never waits for the
// winner-takes-all arithmetic.
template <int NR, bool PARTIAL, bool DBG, int K, int NCH>
// (min blocks = 1: shared memory allows one block per SM anyway, and without it ptxas caps the kernel at 64 registers
// and spills a few; measured C4 final pass 1.070 -> 1.034 ms, C3 0.694 -> 0.688 ms, C1 / C5 unchanged)
__global__ void __launch_bounds__(NR >= 4 ? 256 : 512, 1) aggr_wta_kernel(const __grid_constant__ AggrArgs a) {
  constexpr int DPL = 2 * NR;
  constexpr int NS = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // Placement: ONE block per SM carries up to 5 rows (5 producer + 5 consumer warps), and warp w of
  // a block issues from scheduler w % 4.  The serial SGM chain of a producer is what bounds the
  // kernel, and it stretches when it shares a scheduler with other busy warps (measured with the
  // per-path timeline, tools/trace_aggr.py: 2-warp blocks scattered by the hardware left one row per
  // SM ~35 % slower than the rest).  So the roles are pinned: producers on warps 0..3 (one per
  // scheduler) and, for the fifth row, warp 6 (the scheduler that otherwise hosts only 2 warps);
  // consumers fill the remaining slots.  (Round 2 measured two alternatives on C1, both bit-identical and neither
  // faster: producers on the highest warp ids of their sub-partition, which the arbiter serves first -- 103.0 ->
  // 104.2 us; and shifts / key packing / the blend moved from the alu pipe to IMAD / IMAD.HI on the fma pipe --
  // 103.0 -> 108.3 us, profiles/r02_final_pass_ab.md.)
  const int ppb = blockDim.x >> 6; // paths (rows) per block
  int role, pi;
  if (ppb == 5) {
    const unsigned code = (unsigned)((0xcba4983210ull >> (4 * warp)) & 0xfull); // nibble = role << 3 | row-in-block, for warps 9..0: C4 C3 C2 P4 C1 C0 P3 P2 P1 P0
    role = (int)(code >> 3); pi = (int)(code & 7u);
  } else {
    role = warp >= ppb; pi = role ? warp - ppb : warp;
  }
  const long path = (long)blockIdx.x * ppb + pi;
  const bool valid = path < (long)a.N * a.rows;
  const int D = PARTIAL ? a.D : 64 * NR; // whole lanes: D is a compile-time constant, ring offsets become immediates
  const int PIECE = D * 2;
  const AggrSmem lay = aggr_smem<2, K, NCH>(D);
  unsigned char *wsm = smem_raw + (size_t)pi * lay.total;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(wsm);
  const uint32_t barFull = bar0 + 8 * NCH, barEmpty = barFull + 8 * WTA_TILES;
  if (role == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < NCH + 2 * WTA_TILES; ++i) mbar_init(bar0 + 8 * i, i < NCH ? 1 : 32); // tile hand-over: every lane arrives
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads(); // the only block-level barrier: mbarriers are initialised
  if (!valid) return;
  const PathGeom pg = path_geom(a, path);
  const int steps = pg.steps; // == cols
  const int nact = D / DPL;
  const bool active = !PARTIAL || lane < nact;
  const int loff = lane * NR * 4;
  const int TP = PIECE + 16; // tile row pitch: the per-pixel phase's lanes hit distinct banks
  unsigned char *tile = wsm + lay.tile;
  const int ntiles = (steps + 31) / 32;
  unsigned long long *const tr = g_trace;
  const int tslot = 3 * TRACE_STRIDE + (int)(path % TRACE_STRIDE);
  if (tr && lane == 0 && role == 0) { tr[4 * tslot] = gtimer(); tr[4 * tslot + 2] = smid(); }

  if (role == 0) {
    // =========================== producer: SGM path + blend ===================================
    PathRing<NS, K, NCH> pr;
    pr.bar0 = bar0;
    pr.ring_s = bar0 + lay.ring;
    pr.g[0] = reinterpret_cast<const char *>(a.C + pg.e0);
    pr.g[1] = reinterpret_cast<const char *>(a.aux0 + pg.e0);
    pr.g[2] = nullptr;
    pr.sstride = pg.sstride; pr.steps = steps; pr.PIECE = PIECE; pr.STREAM = NCH * K * PIECE; pr.lane = lane;
    pr.vertical = false; pr.hrev = false; pr.tma = 0; pr.tm = nullptr; pr.c0 = pr.row0 = pr.vrev = 0;
    pr.PX = PIECE; pr.xstride = pg.sstride; pr.c0x = 0;
    const unsigned char *ring = wsm + lay.ring;
    const int STREAM = pr.STREAM;
    const uint32_t selUp = lane == 0 ? 0x5454u : 0x5432u;
    const uint32_t selDn = lane == nact - 1 ? 0x3232u : 0x5432u;
    const int nchunks = (steps + K - 1) / K;
    for (int ci = 0; ci < NCH - 1 && ci < nchunks; ++ci) pr.issue(ci);
    uint32_t L[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) L[r] = active ? 0u : 0xffffffffu;
    char *pDbg0 = DBG ? reinterpret_cast<char *>(a.dbg0 + pg.e0) + loff : nullptr;
    char *pDbg1 = DBG ? reinterpret_cast<char *>(a.dbg1 + pg.e0) + loff : nullptr;

    auto step = [&](const unsigned char *pc, unsigned char *trow) {
      uint32_t c[NR], x0[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) { c[r] = 0; x0[r] = 0; }
      if (active) {
        lds_vec<NR>(pc, c);
        lds_vec<NR>(pc + STREAM, x0);
      }
      sgm_step<NR>(L, c, a.P1P1, a.P2P2, selUp, selDn);
      if (PARTIAL) {
#pragma unroll
        for (int r = 0; r < NR; ++r) L[r] = active ? L[r] : 0xffffffffu;
      }
      // blend: LAll = (L0 + (L1+L2+L3)) / 4 per 16-bit half (aggr.cu:192,222)
      uint32_t la[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) la[r] = ((L[r] + x0[r]) >> 2) & 0x3fff3fffu;
      if (active) {
        st_vec<NR>(trow, la);
        if (DBG) { st_vec<NR>(pDbg0, L); st_vec<NR>(pDbg1, la); }
      }
      if (DBG) { pDbg0 += PIECE; pDbg1 += PIECE; }
    };

    int slot = 0;
    uint32_t parity = 0;
    for (int ci = 0; ci < nchunks; ++ci) {
      const int s0 = ci * K;
      if ((s0 & 31) == 0) { // entering tile t: the consumer must have released it (tile t - WTA_TILES)
        const int t = s0 >> 5;
        if (t >= WTA_TILES) mbar_wait(barEmpty + 8 * (t % WTA_TILES), (uint32_t)((t / WTA_TILES - 1) & 1));
      }
      mbar_wait(bar0 + 8 * slot, parity);
      if (ci + NCH - 1 < nchunks) pr.issue(ci + NCH - 1);
      const int kc = min(K, steps - s0);
      const unsigned char *pc = ring + slot * K * PIECE + loff;
      unsigned char *trow = tile + (s0 % (32 * WTA_TILES)) * TP + loff;
      if (kc == K) {
#pragma unroll
        for (int k = 0; k < K; ++k) { step(pc, trow); pc += PIECE; trow += TP; }
      } else {
        for (int k = 0; k < kc; ++k) { step(pc, trow); pc += PIECE; trow += TP; }
      }
      const int done = s0 + kc;
      if ((done & 31) == 0 || done == steps) { // tile complete: publish it (every lane releases its own stores)
        const uint32_t bar = barFull + 8 * (((done - 1) >> 5) % WTA_TILES);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      }
      if (++slot == NCH) { slot = 0; parity ^= 1; }
    }
    if (tr && lane == 0) tr[4 * tslot + 1] = gtimer();
    return;
  }

  // ============================== consumer: winner-takes-all ===================================
  uint32_t T[DPL], dconst[NR];
#pragma unroll
  for (int k = 0; k < DPL; ++k) T[k] = 0xffffffffu;
#pragma unroll
  for (int r = 0; r < NR; ++r) dconst[r] = (uint32_t)(lane * DPL + 2 * r) | ((uint32_t)(lane * DPL + 2 * r + 1) << 16);
  uint32_t *gkbuf = reinterpret_cast<uint32_t *>(wsm + lay.gk);
  uint16_t *rbuf = reinterpret_cast<uint16_t *>(wsm + lay.rb);
  const size_t rowpix = ((size_t)pg.n * a.rows + pg.q) * a.cols;
  const int k100u = 100 - a.uniq;
  const uint32_t firstmask = lane == 0 ? 0xffffffffu : 0u;
  const bool last_lane = lane == nact - 1;

  auto cstep = [&](const uint32_t (&la)[NR], int k) {
    // keys (value << 16 | d): u32 min == lowest value, then lowest d (wta.cu:30-65)
    uint32_t key[DPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      key[2 * r] = active ? __byte_perm(la[r], dconst[r], 0x1054) : 0xffffffffu;
      key[2 * r + 1] = active ? __byte_perm(la[r], dconst[r], 0x3276) : 0xffffffffu;
    }
    uint32_t lk = key[0];
#pragma unroll
    for (int j = 1; j < DPL; ++j) lk = min(lk, key[j]);
    const uint32_t gk = __reduce_min_sync(FULL, lk);
    if (lane == 0) gkbuf[k] = gk;
    // right disparity: T_x(d) = min(T_{x-1}(d-1), key_x(d))  (wta.cu:183,190-198)
    const uint32_t upT = __shfl_up_sync(FULL, T[DPL - 1], 1) | firstmask;
#pragma unroll
    for (int j = DPL - 1; j >= 1; --j) T[j] = min(T[j - 1], key[j]);
    T[0] = min(upT, key[0]);
    if (last_lane) rbuf[k] = (uint16_t)T[DPL - 1]; // pixel s-(D-1), stored by the tile phase
  };
  auto load_row = [&](const unsigned char *trow, uint32_t (&la)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) la[r] = 0;
    if (active) lds_vec<NR>(trow, la);
  };

  // one lane per pixel: uniqueness + sub-pixel for the (up to 32) pixels of the finished tile
  auto tile_phase = [&](unsigned char *tbase, int t0, int cnt) {
    __syncwarp();
    if (lane < cnt) {
      const float disp = wta_pixel(reinterpret_cast<uint16_t *>(tbase + lane * TP), gkbuf[lane], D, k100u);
      a.dispL[rowpix + t0 + lane] = disp;
      const int xr = t0 + lane - (D - 1);
      if (xr >= 0) a.dispR[rowpix + xr] = rbuf[lane];
    }
    __syncwarp();
  };

  int nextseg = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int ts = t % WTA_TILES;
    mbar_wait(barFull + 8 * ts, (uint32_t)((t / WTA_TILES) & 1));
    const int t0 = t * 32;
    const int cnt = min(32, steps - t0);
    unsigned char *tbase = tile + ts * 32 * TP;
    const unsigned char *trow = tbase + loff;
    uint32_t cur[NR], nxt[NR];
    load_row(trow, cur);
    if (cnt == 32) {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) { // the next row's load is in flight while this one is processed
        trow += TP;
        if (k < 31) load_row(trow, nxt); // (the row after the tile belongs to the producer: never touched)
        cstep(cur, k);
#pragma unroll
        for (int r = 0; r < NR; ++r) cur[r] = nxt[r];
      }
    } else {
      for (int k = 0; k < cnt; ++k) {
        cstep(cur, k);
        trow += TP;
        if (k + 1 < cnt) load_row(trow, cur);
      }
    }
    tile_phase(tbase, t0, cnt);
    {
      const uint32_t bar = barEmpty + 8 * ts;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
    if (nextseg < a.nseg && t0 + cnt >= a.seg_end[nextseg]) { // publish: this row is done up to the segment end
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(a.progress + nextseg, 1u);
      ++nextseg;
    }
  }
  if (active) {
    // pixels whose diagonal leaves the image on the right: x' = cols-1-d, d < D-1
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int d = lane * DPL + k;
      const int xp = a.cols - 1 - d;
      if (d < D - 1 && xp >= 0) a.dispR[rowpix + xp] = (uint16_t)(T[k] & 0xffffu);
    }
  }
  if (tr && lane == 0) tr[4 * tslot + 3] = gtimer();
}

// ---- MODE 2, two rows per warp pair (whole power-of-two disparity ranges, production build) -----------------------
// Same algorithm and hand-over as aggr_wta_kernel, but a row path occupies a HALF-warp: lane = (h, l), h = lane >> 4
// picks one of two adjacent image rows, the 16 lanes l own 2*NR = D/16 consecutive disparities each.  What this buys:
//  * every per-step instruction that does not scale with NR (reduction, shuffles, border selectors, barrier / ring
//    bookkeeping, the consumer's key minimum and stores) is shared by two pixels -- at D = 64 a step of the one-row
//    kernel costs 19 + 19 warp instructions per pixel (producer + consumer), which is what bounds the batched
//    low-resolution workload (BASELINE C4), not HBM;
//  * half as many producer / consumer warps for the same rows: 720 rows are 360 pairs, i.e. at most 3 producer and
//    3 consumer warps per SM instead of 5 + 5 on 4 schedulers, so the serial chain that bounds a single frame shares
//    its scheduler with less.
// The per-half minimum is two CREDUX (halfwarp_min); the shuffles stay full-warp, the lanes at the half borders
// ignore what they receive through the same PRMT selectors that implement the d = 0 / d = D-1 rule.
// A tile is 16 steps x 2 rows = 32 pixels, so the one-lane-per-pixel phase is unchanged.
constexpr int W2_TS = 16; // steps per tile
template <int K, int NCH> __host__ __device__ inline AggrSmem aggr2_smem(int D) {
  AggrSmem s;
  const int piece = D * 2;
  s.ring = 128; // mbarriers: NCH bulk-copy barriers, then WTA_TILES full + WTA_TILES empty
  s.tile = s.ring + 2 /*streams*/ * NCH * 2 /*rows*/ * K * piece;
  s.gk = s.tile + (WTA_TILES * 32 + 1) * (piece + 16);
  s.rb = s.gk + 32 * 4;
  s.total = s.rb + 32 * 2;
  s.total = (s.total + 127) & ~127;
  return s;
}

template <int NR, int K, int NCH>
__global__ void __launch_bounds__(NR >= 8 ? 256 : 512, 1) aggr_wta2_kernel(const __grid_constant__ AggrArgs a) {
  static_assert(W2_TS % K == 0, "a tile is a whole number of chunks");
  constexpr int DPL = 2 * NR;
  constexpr int D = 32 * NR;
  constexpr int PIECE = D * 2;
  constexpr int TP = PIECE + 16;       // tile row pitch: the per-pixel phase's lanes hit distinct banks
  constexpr int ROWCH = K * PIECE;     // bytes of one row's chunk
  constexpr int STREAM = NCH * 2 * ROWCH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, l = lane & 15;
  const bool upper = h != 0;
  // Roles.  ppb pairs per block: producers on warps 0..ppb-1; with three pairs (the single-frame regime: one block
  // per SM) the consumers take warps 3, 4, 5, i.e. schedulers 3, 0, 1 -- one consumer has a scheduler to itself, two
  // share with a producer, one producer runs alone.
  const int ppb = blockDim.x >> 6;
  const int role = warp >= ppb;
  const int pi = role ? warp - ppb : warp;
  const long npaths = (long)a.N * a.rows;
  const long pair = (long)blockIdx.x * ppb + pi;
  const bool valid = 2 * pair < npaths;
  const AggrSmem lay = aggr2_smem<K, NCH>(D);
  unsigned char *wsm = smem_raw + (size_t)pi * lay.total;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(wsm);
  const uint32_t barFull = bar0 + 8 * NCH, barEmpty = barFull + 8 * WTA_TILES;
  if (role == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < NCH + 2 * WTA_TILES; ++i) mbar_init(bar0 + 8 * i, i < NCH ? 1 : 32); // tile hand-over: every lane arrives
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads(); // the only block-level barrier: mbarriers are initialised
  if (!valid) return;
  const long mypath = 2 * pair + h;
  const bool rowvalid = mypath < npaths;              // an odd number of rows: the last pair's upper half repeats the last row
  const PathGeom pg = path_geom(a, rowvalid ? mypath : npaths - 1);
  const int steps = pg.steps; // == cols
  const int loff = l * NR * 4;
  unsigned char *tile = wsm + lay.tile;
  const int ntiles = (steps + W2_TS - 1) / W2_TS;
  unsigned long long *const tr = g_trace;
  const int tslot = 3 * TRACE_STRIDE + (int)((2 * pair) % TRACE_STRIDE);
  if (tr && lane == 0 && role == 0) { tr[4 * tslot] = gtimer(); tr[4 * tslot + 2] = smid(); }

  if (role == 0) {
    // =========================== producer: two SGM paths + blend ==============================
    const uint32_t ring_s = bar0 + lay.ring;
    const unsigned char *ring = wsm + lay.ring;
    // lane 0 issues the copies of both rows: it needs the other row's base as well
    const PathGeom pgB = path_geom(a, 2 * pair + 1 < npaths ? 2 * pair + 1 : npaths - 1);
    const char *gC[2] = {reinterpret_cast<const char *>(a.C + pg.e0), reinterpret_cast<const char *>(a.C + pgB.e0)};
    const char *gS[2] = {reinterpret_cast<const char *>(a.aux0 + pg.e0), reinterpret_cast<const char *>(a.aux0 + pgB.e0)};
    const int nchunks = (steps + K - 1) / K;
    auto issue = [&](int ci) { // chunk ci -> slot ci % NCH: 2 streams x 2 rows, one bulk copy each; all lanes call
      __syncwarp(); // every lane has finished reading the slot that is refilled (the one consumed a chunk ago)
      if (lane != 0) return;
      const int slot = ci % NCH;
      const int s0 = ci * K;
      const int kc = min(K, steps - s0);
      const uint32_t bar = bar0 + 8 * slot;
      const uint32_t dst = ring_s + (uint32_t)(slot * 2 * ROWCH);
      mbar_expect_tx(bar, (uint32_t)(4 * kc * PIECE));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        bulk_g2s(dst + r * ROWCH, gC[r] + (long)s0 * PIECE, (uint32_t)(kc * PIECE), bar);
        bulk_g2s(dst + STREAM + r * ROWCH, gS[r] + (long)s0 * PIECE, (uint32_t)(kc * PIECE), bar);
      }
    };
    for (int ci = 0; ci < NCH - 1 && ci < nchunks; ++ci) issue(ci);
    const uint32_t selUp = l == 0 ? 0x5454u : 0x5432u;
    const uint32_t selDn = l == 15 ? 0x3232u : 0x5432u;
    uint32_t L[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) L[r] = 0u;

    auto step = [&](const unsigned char *pc, unsigned char *trow) {
      uint32_t c[NR], x0[NR];
      lds_vec<NR>(pc, c);
      lds_vec<NR>(pc + STREAM, x0);
      sgm_step<NR, true>(L, c, a.P1P1, a.P2P2, selUp, selDn, upper);
      // blend: LAll = (L0 + (L1+L2+L3)) / 4 per 16-bit half (aggr.cu:192,222)
      uint32_t la[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) la[r] = ((L[r] + x0[r]) >> 2) & 0x3fff3fffu;
      st_vec<NR>(trow, la);
    };

    int slot = 0;
    uint32_t parity = 0;
    for (int ci = 0; ci < nchunks; ++ci) {
      const int s0 = ci * K;
      if ((s0 % W2_TS) == 0) { // entering tile t: the consumer must have released it (tile t - WTA_TILES)
        const int t = s0 / W2_TS;
        if (t >= WTA_TILES) mbar_wait(barEmpty + 8 * (t % WTA_TILES), (uint32_t)((t / WTA_TILES - 1) & 1));
      }
      mbar_wait(bar0 + 8 * slot, parity);
      if (ci + NCH - 1 < nchunks) issue(ci + NCH - 1);
      const int kc = min(K, steps - s0);
      const unsigned char *pc = ring + slot * 2 * ROWCH + h * ROWCH + loff;
      // pixel (row h, step s) of tile slot ts lies at tile row ts*32 + h*16 + s % 16
      unsigned char *trow = tile + (((s0 / W2_TS) % WTA_TILES) * 32 + h * W2_TS + (s0 % W2_TS)) * TP + loff;
      if (kc == K) {
#pragma unroll
        for (int k = 0; k < K; ++k) { step(pc, trow); pc += PIECE; trow += TP; }
      } else {
        for (int k = 0; k < kc; ++k) { step(pc, trow); pc += PIECE; trow += TP; }
      }
      const int done = s0 + kc;
      if ((done % W2_TS) == 0 || done == steps) { // tile complete: publish it (every lane releases its own stores)
        const uint32_t bar = barFull + 8 * (((done - 1) / W2_TS) % WTA_TILES);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      }
      if (++slot == NCH) { slot = 0; parity ^= 1; }
    }
    if (tr && lane == 0) tr[4 * tslot + 1] = gtimer();
    return;
  }

  // ============================== consumer: winner-takes-all ===================================
  uint32_t T[DPL], dconst[NR];
#pragma unroll
  for (int k = 0; k < DPL; ++k) T[k] = 0xffffffffu;
#pragma unroll
  for (int r = 0; r < NR; ++r) dconst[r] = (uint32_t)(l * DPL + 2 * r) | ((uint32_t)(l * DPL + 2 * r + 1) << 16);
  uint32_t *gkbuf = reinterpret_cast<uint32_t *>(wsm + lay.gk);
  uint16_t *rbuf = reinterpret_cast<uint16_t *>(wsm + lay.rb);
  const size_t rowpix = ((size_t)pg.n * a.rows + pg.q) * a.cols;
  const int k100u = 100 - a.uniq;
  const uint32_t firstmask = l == 0 ? 0xffffffffu : 0u;

  auto cstep = [&](const uint32_t (&la)[NR], int k) {
    // keys (value << 16 | d): u32 min == lowest value, then lowest d (wta.cu:30-65)
    uint32_t key[DPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      key[2 * r] = __byte_perm(la[r], dconst[r], 0x1054);
      key[2 * r + 1] = __byte_perm(la[r], dconst[r], 0x3276);
    }
    uint32_t lk = key[0];
#pragma unroll
    for (int j = 1; j < DPL; ++j) lk = min(lk, key[j]);
    const uint32_t gk = halfwarp_min(lk, upper);
    if (l == 0) gkbuf[h * W2_TS + k] = gk;
    // right disparity: T_x(d) = min(T_{x-1}(d-1), key_x(d))  (wta.cu:183,190-198)
    const uint32_t upT = __shfl_up_sync(FULL, T[DPL - 1], 1) | firstmask;
#pragma unroll
    for (int j = DPL - 1; j >= 1; --j) T[j] = min(T[j - 1], key[j]);
    T[0] = min(upT, key[0]);
    if (l == 15) rbuf[h * W2_TS + k] = (uint16_t)T[DPL - 1]; // pixel s-(D-1), stored by the tile phase
  };

  int nextseg = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int ts = t % WTA_TILES;
    mbar_wait(barFull + 8 * ts, (uint32_t)((t / WTA_TILES) & 1));
    const int t0 = t * W2_TS;
    const int cnt = min(W2_TS, steps - t0);
    unsigned char *tbase = tile + ts * 32 * TP;
    const unsigned char *trow = tbase + h * W2_TS * TP + loff;
    uint32_t cur[NR], nxt[NR];
    lds_vec<NR>(trow, cur);
    if (cnt == W2_TS) {
#pragma unroll
      for (int k = 0; k < W2_TS; ++k) { // the next row's load is in flight while this one is processed
        trow += TP;
        if (k < W2_TS - 1) lds_vec<NR>(trow, nxt);
        cstep(cur, k);
#pragma unroll
        for (int r = 0; r < NR; ++r) cur[r] = nxt[r];
      }
    } else {
      for (int k = 0; k < cnt; ++k) {
        cstep(cur, k);
        trow += TP;
        if (k + 1 < cnt) lds_vec<NR>(trow, cur);
      }
    }
    // one lane per pixel: lane (h, l) finishes pixel t0 + l of row h
    __syncwarp();
    if (l < cnt) {
      const float disp = wta_pixel(reinterpret_cast<uint16_t *>(tbase + lane * TP), gkbuf[lane], D, k100u);
      if (rowvalid) {
        a.dispL[rowpix + t0 + l] = disp;
        const int xr = t0 + l - (D - 1);
        if (xr >= 0) a.dispR[rowpix + xr] = rbuf[lane];
      }
    }
    __syncwarp();
    {
      const uint32_t bar = barEmpty + 8 * ts;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
    if (nextseg < a.nseg && t0 + cnt >= a.seg_end[nextseg]) { // publish: these rows are done up to the segment end
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(a.progress + nextseg, 2 * pair + 1 < npaths ? 2u : 1u);
      ++nextseg;
    }
  }
  if (rowvalid) {
    // pixels whose diagonal leaves the image on the right: x' = cols-1-d, d < D-1
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int d = l * DPL + k;
      const int xp = a.cols - 1 - d;
      if (d < D - 1 && xp >= 0) a.dispR[rowpix + xp] = (uint16_t)(T[k] & 0xffffu);
    }
  }
  if (tr && lane == 0) tr[4 * tslot + 3] = gtimer();
}

static int g_wta_pairs = 1; // experiment switch (ssb_debug_set_wta_pairs)
template <int NR> struct AggrCfg {
  static constexpr int K0 = 32 / NR < 2 ? 2 : 32 / NR;
  template <int MODE> static constexpr int K() { return MODE == 1 ? (K0 >= 4 ? K0 / 2 : K0) : K0; }
  template <int MODE> static constexpr int NCH() { return MODE == 0 ? 4 : 3; }
};

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<TensorMapEncodeTiledFn>(p);
  }();
  return fn;
}
// volume [N][rows][cols][D] u16 as a 2-D tensor [N*rows][cols*D], box = K rows x D elements
static bool make_volume_map(CUtensorMap *tm, const uint16_t *base, int N, int rows, int cols, int D, int K) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc || !base) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols * D, (cuuint64_t)N * rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * D * 2};
  const cuuint32_t box[2] = {(cuuint32_t)D, (cuuint32_t)K};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 12-bit packed volume [N][rows][cols][1.5 D bytes] as a 2-D u16 tensor [N*rows][cols * 0.75 D], box = K rows x 0.75 D
static bool make_volume_map12(CUtensorMap *tm, const uint16_t *base, int N, int rows, int cols, int D, int K) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc || !base) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols * (D * 3 / 4), (cuuint64_t)N * rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * (D * 3 / 2)};
  const cuuint32_t box[2] = {(cuuint32_t)(D * 3 / 4), (cuuint32_t)K};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NR, int MODE, bool PARTIAL, bool DBG, int NCHO = 0, bool PACK = false>
static cudaError_t launch_one(const AggrArgs &a_in, cudaStream_t st) {
  AggrArgs a = a_in;
  a.tma = 0;
  constexpr int K = AggrCfg<NR>::template K<MODE>();
  constexpr int NCH = NCHO ? NCHO : AggrCfg<NR>::template NCH<MODE>();
  const size_t smem = (size_t)aggr_smem<MODE, K, NCH>(a.D).total; // per path
  const long npaths = (long)a.N * (a.vertical ? a.cols : a.rows);
  cudaError_t e;
  if constexpr (MODE == 2) {
    static_assert(!PACK, "the final pass reads u16 volumes");
    // Two rows per warp pair (aggr_wta2_kernel) where it measured faster: D = 64 in the throughput regime (C4, 256 envs:
    // final pass 1.03 -> 0.92 ms, 58.7 -> 60.3 k env-frames/s).  It is bit-identical at D = 128 / 256 too, but slower there
    // (C1 103 -> 137 us, C5 0.885 -> 1.12 ms: the per-register work dominates a step from NR = 2 on, and the two CREDUX of
    // the per-half minimum lengthen the serial chain); the switch value 2 keeps those reachable for experiments.
    if constexpr (!DBG && ((!PARTIAL && NR <= 4) || (PARTIAL && NR == 2))) {
      // (D = 96 runs the one-row kernel with 24 of 32 lanes; on half-warps it is 16 whole lanes of 6 disparities)
      const bool d96 = PARTIAL && a.D == 96;
      if (PARTIAL ? (d96 && g_wta_pairs >= 1 && NCHO == 2) : ((g_wta_pairs == 1 && NR == 1 && NCHO == 2) || g_wta_pairs == 2)) {
        constexpr int NR2 = PARTIAL ? 3 : 2 * NR;
        constexpr int K2 = NR2 <= 4 ? 16 : 8;
        constexpr int NCH2 = NCHO == 2 ? 2 : 3;
        auto k2 = aggr_wta2_kernel<NR2, K2, NCH2>;
        const size_t smem2 = (size_t)aggr2_smem<K2, NCH2>(a.D).total; // per pair of rows
        const long npairs = (npaths + 1) / 2;
        long ppb = (npairs + a.nsm - 1) / a.nsm;
        const long fit = (long)((227 * 1024) / smem2);
        const long cap = NR2 >= 8 ? 4 : (NCHO == 2 ? 8 : 3); // (launch bounds: 256 threads at D = 256)
        if (ppb > cap) ppb = cap;
        if (ppb > fit) ppb = fit;
        if (ppb < 1) ppb = 1;
        const size_t bsmem = smem2 * (size_t)ppb;
        if (bsmem > 48 * 1024 && (e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem)) != cudaSuccess) return e;
        k2<<<(unsigned)((npairs + ppb - 1) / ppb), (unsigned)(64 * ppb), bsmem, st>>>(a);
        return cudaGetLastError();
      }
    }
    auto k = aggr_wta_kernel<NR, PARTIAL, DBG, K, NCH>;
    // rows per block: enough for one resident wave with one block per SM when shared memory allows
    // (<= 5 rows in the latency-bound single-frame regime; up to 8 with the 2-slot rings of the
    // batched regime, where rows are plentiful and more resident rows hide more latency)
    long ppb = (npaths + a.nsm - 1) / a.nsm;
    const long fit = (long)((227 * 1024) / smem);
    const long cap = NR >= 4 ? 4 : (NCHO == 2 ? 8 : 5); // (NR >= 4, i.e. D > 128: at most 3 rows fit anyway; launch bounds 256 threads, no 128-register cap)
    if (ppb > cap) ppb = cap;
    if (ppb > fit) ppb = fit;
    if (ppb < 1) ppb = 1;
    const size_t bsmem = smem * (size_t)ppb;
    if (bsmem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem)) != cudaSuccess) return e;
    k<<<(unsigned)((npaths + ppb - 1) / ppb), (unsigned)(64 * ppb), bsmem, st>>>(a);
  } else {
    // (rows of D*2 bytes that are not whole 128-byte lines -- D = 96 -- measured slower as tensor boxes
    // than as bulk pieces: C3 top->bottom 1.61 ms vs 1.74 ms)
    if (a.vertical && a.D <= 256 && a.D * 2 % 128 == 0 && (long)a.N * a.rows < 0x7fffffffL) {
      bool ok = make_volume_map(&a.tm[0], a.C, a.N, a.rows, a.cols, a.D, K);
      if (MODE == 1 && !PACK) ok = ok && make_volume_map(&a.tm[1], a.aux0, a.N, a.rows, a.cols, a.D, K) &&
                                     make_volume_map(&a.tm[2], a.aux1, a.N, a.rows, a.cols, a.D, K);
      if (MODE == 1 && PACK) ok = ok && make_volume_map12(&a.tm[1], a.aux0, a.N, a.rows, a.cols, a.D, K) &&
                                    make_volume_map12(&a.tm[2], a.aux1, a.N, a.rows, a.cols, a.D, K);
      a.tma = ok ? 1 : 0;
    }
    auto k = aggr_kernel<NR, MODE, PARTIAL, DBG, K, NCH, PACK>;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)npaths, 32, smem, st>>>(a);
  }
  return cudaGetLastError();
}

static int nr_for(int D) { return D <= 64 ? 1 : D <= 128 ? 2 : D <= 256 ? 4 : D <= 512 ? 8 : 16; }

template <int MODE, int NCHO = 0> static cudaError_t dispatch(const AggrArgs &a, cudaStream_t st) {
  const int nr = nr_for(a.D);
  const bool partial = a.D != 64 * nr;
  const bool dbg = MODE != 0 && a.dbg0 != nullptr;
  const bool pack = MODE != 2 && a.pack12 && !dbg;
#define SSB_CASE(NRV)                                                                             \
  case NRV:                                                                                       \
    if constexpr (MODE == 0) {                                                                    \
      if (pack) return partial ? launch_one<NRV, 0, true, false, NCHO, true>(a, st) : launch_one<NRV, 0, false, false, NCHO, true>(a, st); \
      return partial ? launch_one<NRV, 0, true, false, NCHO>(a, st) : launch_one<NRV, 0, false, false, NCHO>(a, st); \
    } else if constexpr (MODE == 1) {                                                             \
      if (dbg) return partial ? launch_one<NRV, 1, true, true>(a, st) : launch_one<NRV, 1, false, true>(a, st); \
      if (pack) return partial ? launch_one<NRV, 1, true, false, 0, true>(a, st) : launch_one<NRV, 1, false, false, 0, true>(a, st); \
      return partial ? launch_one<NRV, 1, true, false>(a, st) : launch_one<NRV, 1, false, false>(a, st); \
    } else {                                                                                      \
      if (dbg) return partial ? launch_one<NRV, 2, true, true>(a, st) : launch_one<NRV, 2, false, true>(a, st); \
      return partial ? launch_one<NRV, 2, true, false, NCHO>(a, st) : launch_one<NRV, 2, false, false, NCHO>(a, st); \
    }
  switch (nr) {
    SSB_CASE(1) SSB_CASE(2) SSB_CASE(4) SSB_CASE(8) SSB_CASE(16)
  }
#undef SSB_CASE
  return cudaErrorInvalidValue;
}

} // namespace ssb
// debug hook (not part of include/ss_b200.h): register a device buffer of 4 * 4 * TRACE_STRIDE u64
extern "C" void ssb_debug_set_wta_pairs(int on) { ssb::g_wta_pairs = on; }
extern "C" int ssb_debug_set_aggr_trace(void *device_buffer) {
  return (int)cudaMemcpyToSymbol(ssb::g_trace, &device_buffer, sizeof(device_buffer));
}
namespace ssb {

bool aggr_fast_supported(int D, int cmax, int P1, int P2) {
  if (D < 8 || D > 1024) return false;
  if (D % 8 != 0) return false;              // 16-byte bulk-copy granularity of one pixel's D u16 costs
  if (D % (2 * nr_for(D)) != 0) return false; // whole lanes
  if (P1 < 0 || P2 < 0) return false;
  return 4L * ((long)cmax + P2) <= 65535L && (long)cmax + P2 + P1 <= 65535L;
}

bool aggr_pack12_supported(int D, int cmax, int P2) {
  // every path cost is at most cmax + P2 (aggr.cu:39-76: L <= C + P2).  A packed pixel (1.5 D bytes) must be a whole
  // number of 32-byte sectors, else the vertical passes touch sectors they do not own: D = 96 (144-byte pixels)
  // measured 8.5 % SLOWER packed (C3, 16 envs: 3.04 -> 3.32 ms), D = 64 / 256 4-5 % faster (C4 55.1 -> 57.4 k
  // env-frames/s, C5 467 -> 492 frames/s), D = 128 neutral at two lanes (profiles/r02_pack12_ab.md).
  return D % 64 == 0 && (long)cmax + P2 < 4096;
}

static cudaError_t common_args(AggrArgs &a, const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2, int uniq) {
  if ((long)N * (rows > cols ? rows : cols) > 0x7fffffffL) return cudaErrorInvalidValue;
  a.C = b.C;
  a.N = N; a.rows = rows; a.cols = cols; a.D = D;
  a.P1P1 = (uint32_t)P1 * 0x10001u;
  a.P2P2 = (uint32_t)P2 * 0x10001u;
  a.uniq = uniq;
  a.nsm = sm_count();
  a.pack12 = b.pack12;
  return cudaSuccess;
}

cudaError_t launch_aggr_passes(const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2,
                               int uniq, cudaStream_t stream, cudaStream_t s_aux, cudaEvent_t *ev,
                               const AggrMarks *marks) {
  auto mark = [&](const char *name) { if (marks) marks->mark(marks->ctx, name); };
  AggrArgs a{};
  cudaError_t err;
  if ((err = common_args(a, b, N, rows, cols, D, P1, P2, uniq)) != cudaSuccess) return err;
  // right->left and top->bottom are independent.  Alone, each leaves HBM bandwidth unused (the
  // horizontal pass is bound by its 1280-step serial chain, 92 us + 92 us back to back); forked onto
  // two streams -- with the blocks of BOTH kernels resident together -- the pair takes 164 us.
  AggrArgs h = a;
  h.vertical = 0; h.reverse = 1; h.out = b.L1;
  AggrArgs v = a;
  v.vertical = 1; v.reverse = 0; v.out = b.L2;
  if ((err = cudaEventRecord(ev[0], stream)) != cudaSuccess) return err;
  if ((err = cudaStreamWaitEvent(s_aux, ev[0], 0)) != cudaSuccess) return err;
  // ring depths measured on C1 (horizontal/vertical slots -> pair + following pass, us): 2/2 166+148, 3/2 166+140,
  // 4/2 164+140, 2/3 172+148, 3/3 175+150, 4/3 177+146.  The latency-bound horizontal pass wants the deep
  // ring; the vertical one (tensor copies) finishes last, which leaves its bottom rows in L2 for the
  // bottom->top pass that starts there.
  if ((err = dispatch<0, 2>(v, s_aux)) != cudaSuccess) return err;
  if ((err = dispatch<0, 4>(h, stream)) != cudaSuccess) return err;
  if ((err = cudaEventRecord(ev[1], s_aux)) != cudaSuccess) return err;
  if ((err = cudaStreamWaitEvent(stream, ev[1], 0)) != cudaSuccess) return err;
  mark("aggr_left_down"); // one interval: the two kernels run concurrently
  // bottom->top, accumulating L1+L2+L3
  AggrArgs u = a;
  u.vertical = 1; u.reverse = 1; u.aux0 = b.L1; u.aux1 = b.L2; u.out = b.S3; u.dbg0 = b.dbgL3;
  if ((err = dispatch<1>(u, stream)) != cudaSuccess) return err;
  mark("aggr_up");
  return cudaSuccess;
}

// left->right + blend + winner-takes-all; optional progress counters (see AggrArgs)
cudaError_t launch_aggr_final(const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2,
                              int uniq, cudaStream_t stream, uint32_t *progress, int nseg, const int *seg_end) {
  if (nseg < 0 || nseg > 6 || (nseg && (!progress || !seg_end))) return cudaErrorInvalidValue;
  AggrArgs w{};
  cudaError_t err;
  if ((err = common_args(w, b, N, rows, cols, D, P1, P2, uniq)) != cudaSuccess) return err;
  w.vertical = 0; w.reverse = 0; w.aux0 = b.S3; w.dbg0 = b.dbgL0; w.dbg1 = b.dbgLAll;
  w.dispL = b.dispL; w.dispR = b.dispR;
  w.progress = progress; w.nseg = nseg;
  for (int i = 0; i < nseg; ++i) {
    if ((seg_end[i] & 31) || seg_end[i] <= (i ? seg_end[i - 1] : 0) || seg_end[i] >= cols) return cudaErrorInvalidValue;
    w.seg_end[i] = seg_end[i];
  }
  // more rows than one wave of 5-row blocks can hold: throughput regime (see launch_one)
  if ((long)N * rows > 5L * w.nsm) return dispatch<2, 2>(w, stream);
  return dispatch<2>(w, stream);
}

} // namespace ssb
