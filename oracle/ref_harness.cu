// ref_harness.cu -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim around the UNMODIFIED reference simsense::DepthSensorEngine
// (/root/reference/3rd_party/simsense/include/simsense/core.h:31-104).  oracle/Makefile compiles
// the reference's own .cu files where they lie together with this file into
// oracle/_ref/libsimsense_ref.so.  No reference source is copied into this repository.
//
// The subclass only exposes the `protected` device buffers (core.h:83-93) so that per-stage
// outputs of the real reference can be read back -- the same way the reference's own pybind layer
// subclasses the engine (python/pybind/simsense.cpp:52).  Used by tests (-m gpu), by
// tests/golden/make_golden.py and by bench.py --impl reference.
#include <simsense/core.h>

#include <cstring>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace {

struct RefEngine : public simsense::DepthSensorEngine {
  using simsense::DepthSensorEngine::DepthSensorEngine;

  uint32_t fullRows, fullCols;

  // name -> (device pointer, bytes) for the *currently matched* image size (sz pixels).
  bool stage(const std::string &n, size_t sz, void **p, size_t *bytes) {
    const size_t v = sz * maxDisp;
    const size_t fsz = (size_t)fullRows * fullCols;
    if (n == "rawim0") { *p = d_rawim0; *bytes = fsz; return true; }
    if (n == "rawim1") { *p = d_rawim1; *bytes = fsz; return true; }
    if (n == "noisyim0" && speckleShape > 0) { *p = d_noisyim0; *bytes = fsz; return true; }
    if (n == "noisyim1" && speckleShape > 0) { *p = d_noisyim1; *bytes = fsz; return true; }
    if (n == "recim0" && !rectified) { *p = d_recim0; *bytes = fsz; return true; }
    if (n == "recim1" && !rectified) { *p = d_recim1; *bytes = fsz; return true; }
    if (n == "bboxim0") { *p = d_bboxim0; *bytes = sz; return true; }
    if (n == "bboxim1") { *p = d_bboxim1; *bytes = sz; return true; }
    if (n == "census0") { *p = d_census0; *bytes = 4 * sz; return true; }
    if (n == "census1") { *p = d_census1; *bytes = 4 * sz; return true; }
    if (n == "rawcost" && bfWidth * bfHeight != 1) { *p = d_rawcost; *bytes = 2 * v; return true; }
    if (n == "hsum" && bfWidth * bfHeight != 1) { *p = d_hsum; *bytes = 2 * v; return true; }
    if (n == "cost") { *p = d_cost; *bytes = 2 * v; return true; }
    if (n == "L0") { *p = d_L0; *bytes = 2 * v; return true; }
    if (n == "L1") { *p = d_L1; *bytes = 2 * v; return true; }
    if (n == "L2") { *p = d_L2; *bytes = 2 * v; return true; }
    if (n == "LAll") { *p = d_LAll; *bytes = 2 * v; return true; }
    if (n == "leftDisp") { *p = d_leftDisp; *bytes = 4 * sz; return true; }
    if (n == "rightDisp") { *p = d_rightDisp; *bytes = 2 * sz; return true; }
    if (n == "filteredDisp" && mfSize != 1) { *p = d_filteredDisp; *bytes = 4 * sz; return true; }
    if (n == "bboxDisp") { *p = d_bboxDisp; *bytes = 4 * fsz; return true; }
    if (n == "depth") { *p = d_depth; *bytes = 4 * fsz; return true; }
    if (n == "rgbDepth") { *p = d_rgbDepth; *bytes = 4 * (size_t)rgbRows * rgbCols; return true; }
    return false;
  }
  void zeroBboxDisp() { cudaMemset(d_bboxDisp, 0, 4 * (size_t)fullRows * fullCols); }
};

thread_local std::string g_err;

} // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

void *ref_create(uint32_t rows, uint32_t cols, uint32_t rgbRows, uint32_t rgbCols, float focal,
                 float baseline, float minDepth, float maxDepth, uint64_t seed, float speckleShape,
                 float speckleScale, float mu, float sigma, int rectified, int cw, int ch,
                 uint32_t maxDisp, int bfw, int bfh, int p1, int p2, int uniq, int lr, int mf,
                 float *mapLx, float *mapLy, float *mapRx, float *mapRy, float *a1, float *a2,
                 float *a3, float b1, float b2, float b3, int dilation, float fx, float fy,
                 float skew, float cx, float cy) {
  using simsense::Mat2d;
  try {
    auto *e = new RefEngine(
        rows, cols, rgbRows, rgbCols, focal, baseline, minDepth, maxDepth, seed, speckleShape,
        speckleScale, mu, sigma, rectified != 0, (uint8_t)cw, (uint8_t)ch, maxDisp, (uint8_t)bfw,
        (uint8_t)bfh, (uint8_t)p1, (uint8_t)p2, (uint8_t)uniq, (uint8_t)lr, (uint8_t)mf,
        Mat2d<float>(rows, cols, mapLx), Mat2d<float>(rows, cols, mapLy),
        Mat2d<float>(rows, cols, mapRx), Mat2d<float>(rows, cols, mapRy),
        Mat2d<float>(rows, cols, a1), Mat2d<float>(rows, cols, a2), Mat2d<float>(rows, cols, a3),
        b1, b2, b3, dilation != 0, fx, fy, skew, cx, cy);
    e->fullRows = rows;
    e->fullCols = cols;
    e->zeroBboxDisp(); // neutralise the never-cleared ROI canvas (core.cu:283)
    cudaDeviceSynchronize();
    return e;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return nullptr;
  }
}

void ref_destroy(void *h) { delete static_cast<RefEngine *>(h); }

int ref_compute_host(void *h, uint8_t *left, uint8_t *right, int bbox, uint32_t x, uint32_t y,
                     uint32_t w, uint32_t hgt) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    e->compute(simsense::Mat2d<uint8_t>(e->fullRows, e->fullCols, left),
               simsense::Mat2d<uint8_t>(e->fullRows, e->fullCols, right), bbox != 0, x, y, w, hgt);
    return 0;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return 1;
  }
}

int ref_compute_device(void *h, void *leftRgba, void *rightRgba, int bbox, uint32_t x, uint32_t y,
                       uint32_t w, uint32_t hgt) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    e->compute(leftRgba, rightRgba, bbox != 0, x, y, w, hgt);
    return 0;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return 1;
  }
}

// Copies getMat2d() (core.cu:347-362) into out; returns rows*cols or -1.
long ref_get_depth(void *h, float *out) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    auto m = e->getMat2d();
    std::memcpy(out, m.data(), sizeof(float) * m.rows() * m.cols());
    return (long)(m.rows() * m.cols());
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

void *ref_get_depth_device(void *h) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    return e->getCudaPtr();
  } catch (std::exception &ex) {
    g_err = ex.what();
    return nullptr;
  }
}

long ref_get_point_cloud(void *h, float *out) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    auto m = e->getPointCloudMat2d();
    std::memcpy(out, m.data(), sizeof(float) * m.rows() * m.cols());
    return (long)m.rows();
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

long ref_get_rgb_point_cloud(void *h, void *rgbaDevice, float *out) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    auto m = e->getRgbPointCloudMat2d(rgbaDevice);
    std::memcpy(out, m.data(), sizeof(float) * m.rows() * m.cols());
    return (long)m.rows();
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// getRgbPointCloudCudaPtr() (core.cu:436-455): the device-side getter, no host copy.  Returns the pointer or null.
void *ref_get_rgb_point_cloud_device(void *h, void *rgbaDevice) {
  auto *e = static_cast<RefEngine *>(h);
  try {
    return e->getRgbPointCloudCudaPtr(rgbaDevice);
  } catch (std::exception &ex) {
    g_err = ex.what();
    return nullptr;
  }
}

// Reads a protected stage buffer back to host.  matchedPixels = rows*cols of the image that was
// actually matched (ROI size when bbox was used).  Returns bytes copied or -1.
long ref_get_stage(void *h, const char *name, size_t matchedPixels, void *out, size_t capacity) {
  auto *e = static_cast<RefEngine *>(h);
  void *p = nullptr;
  size_t bytes = 0;
  if (!e->stage(name, matchedPixels, &p, &bytes)) {
    g_err = std::string("unknown or unallocated stage: ") + name;
    return -1;
  }
  if (bytes > capacity) {
    g_err = "capacity too small";
    return -1;
  }
  if (cudaMemcpy(out, p, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
    g_err = "cudaMemcpy failed";
    return -1;
  }
  return (long)bytes;
}

void ref_set_penalties(void *h, int p1, int p2) {
  static_cast<RefEngine *>(h)->setPenalties((uint8_t)p1, (uint8_t)p2);
}
void ref_set_census_window_size(void *h, int w, int hgt) {
  static_cast<RefEngine *>(h)->setCensusWindowSize((uint8_t)w, (uint8_t)hgt);
}
void ref_set_matching_block_size(void *h, int w, int hgt) {
  static_cast<RefEngine *>(h)->setMatchingBlockSize((uint8_t)w, (uint8_t)hgt);
}
void ref_set_uniqueness_ratio(void *h, int u) {
  static_cast<RefEngine *>(h)->setUniquenessRatio((uint8_t)u);
}
void ref_set_lr_max_diff(void *h, int d) { static_cast<RefEngine *>(h)->setLrMaxDiff((uint8_t)d); }

} // extern "C"
