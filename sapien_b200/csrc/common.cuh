// common.cuh -- shared device helpers for the sm_100a stereo depth kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssb {

constexpr unsigned FULL = 0xffffffffu;

// ---- vector loads/stores of NR packed u16x2 registers (NR = 1,2,4,8,16) -----------------------
template <int NR> struct Vec;
template <> struct Vec<1> {
  static __device__ __forceinline__ void ld(const uint16_t *p, uint32_t (&r)[1]) {
    r[0] = __ldg(reinterpret_cast<const unsigned int *>(p));
  }
  static __device__ __forceinline__ void st(uint16_t *p, const uint32_t (&r)[1]) {
    *reinterpret_cast<unsigned int *>(p) = r[0];
  }
};
template <> struct Vec<2> {
  static __device__ __forceinline__ void ld(const uint16_t *p, uint32_t (&r)[2]) {
    uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    r[0] = v.x; r[1] = v.y;
  }
  static __device__ __forceinline__ void st(uint16_t *p, const uint32_t (&r)[2]) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(r[0], r[1]);
  }
};
template <int NR> struct Vec {
  static_assert(NR % 4 == 0, "NR must be 1, 2 or a multiple of 4");
  static __device__ __forceinline__ void ld(const uint16_t *p, uint32_t (&r)[NR]) {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i) {
      uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + i);
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  }
  static __device__ __forceinline__ void st(uint16_t *p, const uint32_t (&r)[NR]) {
#pragma unroll
    for (int i = 0; i < NR / 4; ++i)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  }
};

__device__ __forceinline__ uint32_t pack2(uint32_t v) { return v * 0x10001u; }
__device__ __forceinline__ uint32_t min_halves(uint32_t v) {
  const uint32_t lo = v & 0xffffu, hi = v >> 16;
  return lo < hi ? lo : hi;
}

// CUDA float->int conversion semantics are what the reference relies on (roundf then (int)).
__device__ __forceinline__ int round_to_int(float v) { return (int)roundf(v); }

} // namespace ssb
