// common.cuh -- shared helpers for the sm_100a stereo depth kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssb {

constexpr unsigned FULL = 0xffffffffu;

// SM count of the CURRENT device (launchers size their grids from it).  Cached per device ordinal:
// one process may drive engines on several, possibly different, devices.
inline int sm_count() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

} // namespace ssb
