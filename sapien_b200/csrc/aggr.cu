// aggr.cu -- SGM path aggregation (4 paths), blend and winner-takes-all for sm_100a.
//
// Replaces the reference kernels aggrLeft2Right / aggrRight2Left / aggrTop2Bottom /
// aggrBottom2Top (3rd_party/simsense/src/aggr.cu:29-230) and winnerTakesAll
// (src/wta.cu:170-214).  Semantics reproduced exactly; the
// organisation is new:
//
//  * one WARP walks one path (an image row or column); a lane owns DPL = 2*NR consecutive
//    disparities packed as u16x2 registers, so a path step is ~25 instructions and needs no
//    shared memory, no __syncthreads and no atomics (the reference spends 3 barriers and one
//    shared atomicMin per step with one thread per disparity);
//  * d-1 / d+1 neighbours come from funnel shifts inside the lane plus one shuffle each way;
//    the running minimum is one redux.sync; the min/add chain is Blackwell DPX
//    (__viaddmin_u16x2 -> VIADDMNMX.U16x2, __vminu2 -> VIMNMX.U16x2);
//  * the cost volume is streamed through a per-warp shared-memory ring filled by cp.async
//    (LDGSTS) PF path steps ahead and retired with cp.async.wait_group, so a warp keeps
//    PF*(1+NAUX) independent 256 B (D=128) requests in flight: the walk is latency-hidden and the
//    pass becomes HBM-bound instead of barrier-bound.  (A register prefetch ring does not work:
//    the 6 scoreboard slots alias loads issued 6 steps apart, measured 1 DRAM latency per step.);
//  * pass order is  (right->left || top->bottom)  ->  bottom->top (+L1+L2)  ->  left->right.
//    The last pass holds LAll(y,x,:) = (L0+L1+L2+L3)/4 in registers and does the winner-takes-all
//    in place: left disparity (uniqueness, sub-pixel) per step, right disparity through a
//    register recurrence along the diagonal  T_x(d) = min(T_{x-1}(d-1), LAll(x,d)).  LAll is never
//    written to memory (the reference writes it and launches one block per pixel to read it back).
//
// HBM traffic: 2V + 2V + 4V + 2V = 10 V for aggregation + WTA, against ~13 V in the reference
// (V = one u16 volume).
#include "common.cuh"
#include "kernels.h"

namespace ssb {

struct AggrArgs {
  const uint16_t *C;
  const uint16_t *aux0, *aux1;
  uint16_t *out;  // plain passes: L ; up pass: L+aux0+aux1
  uint16_t *dbg0; // up pass: L3 ; wta pass: L0
  uint16_t *dbg1; // wta pass: LAll
  float *dispL;
  uint16_t *dispR;
  int N, rows, cols, D;
  int vertical, reverse;
  uint32_t P1P1, P2P2, BIG;
  int uniq;
};

template <int NR>
__device__ __forceinline__ void sgm_step(uint32_t (&L)[NR], const uint32_t (&c)[NR], uint32_t P1P1,
                                         uint32_t P2P2, uint32_t BIG, bool first_lane,
                                         bool last_lane, bool active) {
  // running minimum over all disparities of the previous pixel
  uint32_t mn = L[0];
#pragma unroll
  for (int j = 1; j < NR; ++j) mn = __vminu2(mn, L[j]);
  const uint32_t m = __reduce_min_sync(FULL, min_halves(mn));
  const uint32_t mm = pack2(m);
  const uint32_t mP2 = mm + P2P2;
  // d-1 of my first element lives in the previous lane, d+1 of my last in the next lane
  uint32_t up = __shfl_up_sync(FULL, L[NR - 1], 1);
  uint32_t dn = __shfl_down_sync(FULL, L[0], 1);
  if (first_lane) up = BIG << 16;
  if (last_lane) dn = BIG;
  uint32_t nl[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const uint32_t lo = (j == 0) ? up : L[j - 1];
    const uint32_t hi = (j == NR - 1) ? dn : L[j + 1];
    const uint32_t lm1 = __funnelshift_l(lo, L[j], 16); // L(d-1) for both halves
    const uint32_t lp1 = __funnelshift_r(L[j], hi, 16); // L(d+1) for both halves
    uint32_t t = __viaddmin_u16x2(lm1, P1P1, L[j]);
    t = __viaddmin_u16x2(lp1, P1P1, t);
    t = __vminu2(t, mP2);
    nl[j] = t - mm + c[j]; // per-half: t >= m and result < 65536 in the fast regime
  }
  const uint32_t bigbig = pack2(BIG);
#pragma unroll
  for (int j = 0; j < NR; ++j) L[j] = active ? nl[j] : bigbig;
}

// ---- cp.async (LDGSTS) staging: every lane copies its own NR words global -> shared ------------
template <int NR> __device__ __forceinline__ void cp_async_vec(uint32_t saddr, const uint16_t *g) {
  if constexpr (NR == 1) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
  } else if constexpr (NR == 2) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr + 16 * i), "l"(g + 8 * i) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int NR> __device__ __forceinline__ void lds_vec(const uint32_t *p, uint32_t (&r)[NR]) {
  if constexpr (NR == 1) {
    r[0] = p[0];
  } else if constexpr (NR == 2) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    r[0] = v.x; r[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i) {
      const uint4 v = reinterpret_cast<const uint4 *>(p)[i];
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  }
}

// Shared memory per warp: (1+NAUX) streams x (PF+1) slots x 32 lanes x NR words, then (WTA) D u16.
template <int NR, int NAUX, int PF> constexpr size_t ring_bytes_per_warp() {
  return (size_t)(1 + NAUX) * (PF + 1) * 32 * NR * 4;
}

template <int NR, int NAUX, bool WTA, int PF>
__global__ void __launch_bounds__(256) aggr_kernel(const AggrArgs a) {
  constexpr int DPL = 2 * NR;
  constexpr int NSLOT = PF + 1;             // prefetch distance PF, one spare slot (no WAR hazard)
  constexpr int SLOT_WORDS = 32 * NR;
  constexpr int STREAM_WORDS = NSLOT * SLOT_WORDS;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const long path = (long)blockIdx.x * wpb + warp;
  const int per_env = a.vertical ? a.cols : a.rows;
  if (path >= (long)a.N * per_env) return;
  const int n = (int)(path / per_env);
  const int q = (int)(path - (long)n * per_env);
  const int steps = a.vertical ? a.rows : a.cols;
  const size_t D = (size_t)a.D;
  const size_t env = (size_t)n * a.rows * a.cols * D;
  size_t base;
  long stride;
  if (a.vertical) { stride = (long)a.cols * (long)D; base = env + (size_t)q * D; }
  else { stride = (long)D; base = env + (size_t)q * a.cols * D; }
  if (a.reverse) { base += (size_t)(steps - 1) * (size_t)stride; stride = -stride; }
  const int d0 = lane * DPL;
  const bool active = d0 < a.D;
  const bool first_lane = lane == 0;
  const bool last_lane = lane == a.D / DPL - 1;
  base += active ? d0 : 0;
  const uint16_t *pC = a.C + base;
  const uint16_t *pA0 = NAUX > 0 ? a.aux0 + base : nullptr;
  const uint16_t *pA1 = NAUX > 1 ? a.aux1 + base : nullptr;
  const uint32_t bigbig = pack2(a.BIG);

  uint32_t *ring = smem + (size_t)warp * ((1 + NAUX) * STREAM_WORDS) + lane * NR;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  auto issue = [&](int step, int slot) {
    const long off = (long)step * stride;
    cp_async_vec<NR>(ring_s + (uint32_t)(slot * SLOT_WORDS * 4), pC + off);
    if (NAUX > 0) cp_async_vec<NR>(ring_s + (uint32_t)((STREAM_WORDS + slot * SLOT_WORDS) * 4), pA0 + off);
    if (NAUX > 1) cp_async_vec<NR>(ring_s + (uint32_t)((2 * STREAM_WORDS + slot * SLOT_WORDS) * 4), pA1 + off);
  };
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    if (active && j < steps) issue(j, j);
    cp_async_commit();
  }

  uint32_t L[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) L[r] = bigbig;

  // winner-takes-all state (left->right pass only)
  uint32_t T[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) T[k] = 0xffffffffu;
  uint16_t *my_la = reinterpret_cast<uint16_t *>(smem + (size_t)wpb * ((1 + NAUX) * STREAM_WORDS)) + (size_t)warp * a.D;
  const size_t rowpix = WTA ? ((size_t)n * a.rows + q) * a.cols : 0;
  const int k100u = 100 - a.uniq;

  int slot = 0, fill = PF;
  for (int s = 0; s < steps; ++s) {
    cp_async_wait<PF - 1>(); // the group of step s has landed
    uint32_t c[NR], x0[NR], x1[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { c[r] = 0; x0[r] = 0; x1[r] = 0; }
    if (active) {
      lds_vec<NR>(ring + slot * SLOT_WORDS, c);
      if (NAUX > 0) lds_vec<NR>(ring + STREAM_WORDS + slot * SLOT_WORDS, x0);
      if (NAUX > 1) lds_vec<NR>(ring + 2 * STREAM_WORDS + slot * SLOT_WORDS, x1);
      if (s + PF < steps) issue(s + PF, fill); // refills the slot consumed one step ago
    }
    cp_async_commit();
    slot = slot + 1 == NSLOT ? 0 : slot + 1;
    fill = fill + 1 == NSLOT ? 0 : fill + 1;
    if (s == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) L[r] = active ? c[r] : bigbig;
    } else {
      sgm_step<NR>(L, c, a.P1P1, a.P2P2, a.BIG, first_lane, last_lane, active);
    }
    const long off = (long)s * stride;
    if (!WTA) {
      if (active) {
        uint32_t o[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) o[r] = L[r] + (NAUX > 0 ? x0[r] : 0u) + (NAUX > 1 ? x1[r] : 0u);
        Vec<NR>::st(a.out + base + off, o);
        if (NAUX > 0 && a.dbg0) Vec<NR>::st(a.dbg0 + base + off, L);
      }
    } else {
      // ---- blend: LAll = (L0 + (L1+L2+L3)) / 4, per 16-bit half ---------------------------
      uint32_t la[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) la[r] = ((L[r] + x0[r]) >> 2) & 0x3fff3fffu;
      if (active) {
        if (a.dbg0) Vec<NR>::st(a.dbg0 + base + off, L);
        if (a.dbg1) Vec<NR>::st(a.dbg1 + base + off, la);
        Vec<NR>::st(my_la + d0, la);
      }
      // ---- keys (value<<16 | d): u32 min == lowest value, then lowest d -----------------
      uint32_t key[DPL];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        key[2 * r] = active ? ((la[r] << 16) | (uint32_t)(d0 + 2 * r)) : 0xffffffffu;
        key[2 * r + 1] = active ? ((la[r] & 0xffff0000u) | (uint32_t)(d0 + 2 * r + 1)) : 0xffffffffu;
      }
      uint32_t lk = key[0];
#pragma unroll
      for (int k = 1; k < DPL; ++k) lk = min(lk, key[k]);
      const uint32_t gk = __reduce_min_sync(FULL, lk);
      const int mval = (int)(gk >> 16);
      const int dstar = (int)(gk & 0xffffu);
      // ---- uniqueness (wta.cu:203): all d: LAll(d)*(100-u) >= m*100 or |d-d*|<=1 ---------
      bool uniq_ok;
      if (k100u > 0) {
        uint32_t m2 = 0xffffu; // smallest value outside d*-1..d*+1 (the test is monotone)
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          const bool excl = (unsigned)(d0 + k - dstar + 1) <= 2u;
          const uint32_t v = key[k] >> 16;
          m2 = min(m2, excl ? 0xffffu : v);
        }
        m2 = __reduce_min_sync(FULL, m2);
        uniq_ok = (int)m2 * k100u >= mval * 100;
      } else {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          const int v = (int)(key[k] >> 16);
          const int dd = d0 + k - dstar;
          ok = ok && (!active || v * k100u >= mval * 100 || (dd >= -1 && dd <= 1));
        }
        uniq_ok = __all_sync(FULL, ok);
      }
      __syncwarp();
      float disp = (float)dstar;
      if (!uniq_ok) {
        disp = -1.0f;
      } else if (dstar != 0 && dstar != a.D - 1) {
        const int y0 = my_la[dstar - 1], y2 = my_la[dstar + 1];
        const float sub = (float)((1.0 * (double)(y2 - y0)) / (2.0 * (double)(y0 - 2 * mval + y2)));
        disp = (float)dstar - sub;
      }
      __syncwarp();
      if (lane == 0) a.dispL[rowpix + s] = disp;
      // ---- right disparity: T_x(d) = min(T_{x-1}(d-1), key_x(d)) --------------------------
      const uint32_t upT = __shfl_up_sync(FULL, T[DPL - 1], 1);
#pragma unroll
      for (int k = DPL - 1; k >= 1; --k) T[k] = min(T[k - 1], key[k]);
      T[0] = first_lane ? key[0] : min(upT, key[0]);
      if (last_lane && s >= a.D - 1) a.dispR[rowpix + s - (a.D - 1)] = (uint16_t)(T[DPL - 1] & 0xffffu);
    }
  }
  cp_async_wait<0>();
  if (WTA && active) {
    // pixels whose diagonal leaves the image on the right: x' = cols-1-d, d < D-1
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int d = d0 + k;
      const int xp = a.cols - 1 - d;
      if (d < a.D - 1 && xp >= 0) a.dispR[rowpix + xp] = (uint16_t)(T[k] & 0xffffu);
    }
  }
}

constexpr int pf_for(int NR, int NAUX) {
  // prefetch distance in path steps; the ring costs (1+NAUX)*(PF+1)*128*NR bytes per warp
  int v = 96 / (NR * (1 + NAUX));
  return v > 32 ? 32 : (v < 3 ? 3 : v);
}

template <int NR, int NAUX, bool WTA>
static cudaError_t launch_one(const AggrArgs &a, int wpb, cudaStream_t st) {
  constexpr int PF = pf_for(NR, NAUX);
  const long npaths = (long)a.N * (a.vertical ? a.cols : a.rows);
  const unsigned blocks = (unsigned)((npaths + wpb - 1) / wpb);
  const size_t smem = (size_t)wpb * ring_bytes_per_warp<NR, NAUX, PF>() + (WTA ? (size_t)wpb * a.D * sizeof(uint16_t) : 0);
  auto k = aggr_kernel<NR, NAUX, WTA, PF>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<blocks, wpb * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <int NAUX, bool WTA>
static cudaError_t dispatch(const AggrArgs &a, int wpb, cudaStream_t st) {
  const int D = a.D;
  if (D <= 64) return launch_one<1, NAUX, WTA>(a, wpb, st);
  if (D <= 128) return launch_one<2, NAUX, WTA>(a, wpb, st);
  if (D <= 256) return launch_one<4, NAUX, WTA>(a, wpb, st);
  if (D <= 512) return launch_one<8, NAUX, WTA>(a, wpb, st);
  return launch_one<16, NAUX, WTA>(a, wpb, st);
}

static int dpl_for(int D) { return D <= 64 ? 2 : D <= 128 ? 4 : D <= 256 ? 8 : D <= 512 ? 16 : 32; }

bool aggr_fast_supported(int D, int cmax, int P1, int P2) {
  if (D < 2 || D > 1024) return false;
  if (D % dpl_for(D) != 0) return false;
  if (P1 < 0 || P2 < 0) return false;
  return 4L * ((long)cmax + P2) <= 65535L && (long)cmax + P2 + P1 <= 65535L;
}

cudaError_t launch_aggr_wta(const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2,
                            int uniq, cudaStream_t stream, cudaStream_t s_aux, cudaEvent_t *ev,
                            const AggrMarks *marks) {
  auto mark = [&](const char *name) { if (marks) marks->mark(marks->ctx, name); };
  AggrArgs a{};
  a.C = b.C;
  a.N = N; a.rows = rows; a.cols = cols; a.D = D;
  a.P1P1 = (uint32_t)P1 * 0x10001u;
  a.P2P2 = (uint32_t)P2 * 0x10001u;
  a.BIG = 0xffffu - (uint32_t)P1;
  a.uniq = uniq;
  cudaError_t err;
  // fork: right->left on the aux stream, top->bottom on the main stream
  if ((err = cudaEventRecord(ev[0], stream)) != cudaSuccess) return err;
  if ((err = cudaStreamWaitEvent(s_aux, ev[0], 0)) != cudaSuccess) return err;
  AggrArgs h = a;
  h.vertical = 0; h.reverse = 1; h.out = b.L1;
  if ((err = dispatch<0, false>(h, 2, s_aux)) != cudaSuccess) return err;
  if ((err = cudaEventRecord(ev[1], s_aux)) != cudaSuccess) return err;
  AggrArgs v = a;
  v.vertical = 1; v.reverse = 0; v.out = b.L2;
  if ((err = dispatch<0, false>(v, 8, stream)) != cudaSuccess) return err;
  mark("aggr_down");
  if ((err = cudaStreamWaitEvent(stream, ev[1], 0)) != cudaSuccess) return err;
  mark("aggr_left_tail"); // time the right->left pass (aux stream) outlives the top->bottom one
  // bottom->top, accumulating L1+L2+L3
  AggrArgs u = a;
  u.vertical = 1; u.reverse = 1; u.aux0 = b.L1; u.aux1 = b.L2; u.out = b.S3; u.dbg0 = b.dbgL3;
  if ((err = dispatch<2, false>(u, 8, stream)) != cudaSuccess) return err;
  mark("aggr_up");
  // left->right + blend + winner-takes-all
  AggrArgs w = a;
  w.vertical = 0; w.reverse = 0; w.aux0 = b.S3; w.dbg0 = b.dbgL0; w.dbg1 = b.dbgLAll;
  w.dispL = b.dispL; w.dispR = b.dispR;
  err = dispatch<1, true>(w, 2, stream);
  mark("aggr_right_wta");
  return err;
}

} // namespace ssb
