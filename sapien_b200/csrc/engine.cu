// engine.cu -- host side of the B200 stereo depth engine and its C ABI (include/ss_b200.h).
//
// Replaces simsense::DepthSensorEngine's host code (3rd_party/simsense/src/core.cu:65-787):
// buffer ownership, stage sequencing, getters/setters.  Differences by design: bound to an
// explicit device; one stream + one helper stream with event fork/join instead of 3 streams and
// 13-16 cudaDeviceSynchronize per frame; status codes instead of exit(); a batch dimension; 3-4
// resident cost volumes instead of 7.
#include "../../include/ss_b200.h"
#include "kernels.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h> // header-only; ranges cost nothing unless a profiler is attached

using namespace ssb;

namespace {

thread_local std::string g_err;

// debug hooks (not part of include/ss_b200.h): 0 turns the 12-bit storage of the path volumes off (A/B measurements);
// a cap on the environments per wave makes engines created afterwards run their batches in several waves (tests)
int g_allow_pack12 = 1;
int g_max_wave = 0;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(SS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));               \
  } while (0)

// NVTX range around the enqueue of one stage (the reference brackets its host code the same way,
// include/sapien/profiler.h:33-38).  Device time per stage: ss_set_profiling / ss_get_stage_times.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int census_bits(int cw, int ch) {
  const int bits = ((ch - 1) / 2) * cw + cw / 2;
  return std::min(bits, 32);
}

// Ranges of python/py_package/sensor/simsense_component.py:54-134.
int validate_params(const ss_config &c) {
  if (c.rows < 32 || c.cols < 32) return fail(SS_ERR_INVALID, "Infrared resolution (width and height) must be integer and no less than 32");
  if (c.census_width <= 0 || c.census_height <= 0 || c.census_width % 2 == 0 || c.census_height % 2 == 0 || c.census_width * c.census_height > 65)
    return fail(SS_ERR_INVALID, "census_width and census_height must be positive odd integers and their product should be no larger than 65");
  if (c.max_disp < 32 || c.max_disp > 1024) return fail(SS_ERR_INVALID, "max_disp must be integer and within range [32, 1024]");
  if (c.bf_width <= 0 || c.bf_height <= 0 || c.bf_width % 2 == 0 || c.bf_height % 2 == 0 || c.bf_width * c.bf_height > 256)
    return fail(SS_ERR_INVALID, "block_width and block_height must be positive odd integers and their product should be no larger than 256");
  if (c.bf_width / 2 >= (int)c.cols || c.bf_height / 2 >= (int)c.rows) return fail(SS_ERR_INVALID, "matching block larger than the image");
  if (c.p1 <= 0 || c.p2 <= 0 || c.p1 >= c.p2 || c.p2 >= 224)
    return fail(SS_ERR_INVALID, "p1_penalty must be positive integer less than p2_penalty and p2_penalty be positive integer less than 224");
  if (c.uniq_ratio < 0 || c.uniq_ratio > 255) return fail(SS_ERR_INVALID, "uniqueness_ratio must be positive integer and no larger than 255");
  if (c.lr_max_diff < -1 || c.lr_max_diff > 255) return fail(SS_ERR_INVALID, "lr_max_diff must be integer and within the range [0, 255]");
  if (c.mf_size != 1 && c.mf_size != 3 && c.mf_size != 5 && c.mf_size != 7) return fail(SS_ERR_INVALID, "Median filter size choices are 1, 3, 5, 7");
  return SS_OK;
}

} // namespace

struct Core {
  ss_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr, aux = nullptr, cpy = nullptr;
  // host-input frames: uploads on `up` into double-buffered raw images, their front-ends on `fr`, so that the uploads
  // and the front-end of frame k+1 overlap frame k (ss_submit_host_u8 keeps up to two frames in flight)
  cudaStream_t up = nullptr, fr = nullptr;
  cudaEvent_t ev_upL = nullptr, ev_upR = nullptr;
  cudaEvent_t ev_lastband[2] = {nullptr, nullptr}; // column bands (helper stream) of the frame in slot s have finished
  int band_slot = 0;
  cudaEvent_t ev_rawfree[2] = {nullptr, nullptr}; // front-end that read raw slot s has finished
  cudaEvent_t ev_done[2] = {nullptr, nullptr};    // host delivery of the frame in output slot s has finished
  uint64_t slot_ticket[2] = {0, 0};               // ticket (frame number + 1) whose delivery ev_done[s] stands for, 0 = none
  bool async_unwaited = false;                    // the last frame was submitted asynchronously and nobody has waited for it yet
  bool band_pending = false;                      // the last frame's column bands (helper stream) are not yet joined into the main stream
  uint8_t *raw0s[2] = {nullptr, nullptr}, *raw1s[2] = {nullptr, nullptr};
  float *out_pair[2] = {nullptr, nullptr};        // final depth of even / odd frames (a frame's read-back overlaps the next frame)
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaEvent_t ev_dep = nullptr; // ss_wait_stream: extra producer stream(s) the next compute is ordered after
  bool dep_pending = false;
  // banded output (ss_bind_output_host): the final SGM pass runs in column segments and the depth map
  // streams to the bound host buffer band by band while the remaining segments compute
  static constexpr int MAXSEG = 6;
  cudaEvent_t ev_seg[MAXSEG] = {}, ev_band[MAXSEG] = {};
  float *host_out = nullptr;    // caller-bound output (ss_bind_output_host)
  size_t host_cap = 0;
  float *staging = nullptr;     // engine-owned pinned output of host-input frames when nothing is bound (the
                                // reference stages every read-back through an engine-owned host buffer, core.cu:347-362)
  float *stream_dst = nullptr;  // where the last compute streamed its depth map (host_out, staging or null)
  bool streamed = false;        // the last compute delivered its depth map into stream_dst
  std::vector<int> rgb_sufmin;  // [cols+1] smallest RGB column any matched column >= x can splat into (empty: banding off)
  uint32_t *progress = nullptr; // [MAXSEG] rows finished per column segment of the final pass

  std::vector<void *> allocs;
  float *mapLx = nullptr, *mapLy = nullptr, *mapRx = nullptr, *mapRy = nullptr;
  float *a1 = nullptr, *a2 = nullptr, *a3 = nullptr;
  bool has_cal = false; // ss_create_calibrated: no planes, the kernels evaluate them from `cal`
  ss_calibration cal{};
  uint8_t *raw0 = nullptr, *raw1 = nullptr, *im0 = nullptr, *im1 = nullptr;
  uint32_t *cen0 = nullptr, *cen1 = nullptr;
  uint16_t *C = nullptr, *L1 = nullptr, *L2 = nullptr, *S3 = nullptr;
  uint16_t *dbgL0 = nullptr, *dbgL3 = nullptr, *dbgLAll = nullptr;
  float *dispL = nullptr, *disp_lr = nullptr, *disp_med = nullptr, *disp_full = nullptr;
  float *depth = nullptr, *canvas = nullptr, *out = nullptr, *pc = nullptr, *rgbpc = nullptr;
  float *canvas_pair[2] = {nullptr, nullptr}; // splat canvases of even / odd frames (see the front-end launch)
  cudaEvent_t ev_cost = nullptr, ev_front = nullptr; // previous frame past its HBM-bound passes / front-end of this frame done
  uint16_t *dispR = nullptr;
  int wave = 1;
  bool computed = false;
  bool last_pack12 = false; // the last frame stored L1 / L2 12-bit packed (ss_get_stage_host unpacks)
  int mrows = 0, mcols = 0; // matched size of the last compute
  uint64_t frame = 0;
  uint64_t frame_id = 0; // engine-wide frame number (set by the lane dispatcher): keys the IR-noise stream, so that
                         // consecutive frames differ whichever lane they run on
  int launches = 0;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> pev;
  std::vector<const char *> pnames;
  std::vector<float> ptimes;

  size_t fsz() const { return (size_t)cfg.rows * cfg.cols; }
  size_t rsz() const { return cfg.registration ? (size_t)cfg.rgb_rows * cfg.rgb_cols : fsz(); }
  uint32_t out_rows() const { return cfg.registration ? cfg.rgb_rows : cfg.rows; }
  uint32_t out_cols() const { return cfg.registration ? cfg.rgb_cols : cfg.cols; }

  template <class T> int alloc(T **p, size_t count) {
    void *q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    allocs.push_back(q);
    *p = static_cast<T *>(q);
    return SS_OK;
  }
  int ensure_generic_volumes() {
    const size_t v = (size_t)wave * fsz() * cfg.max_disp;
    if (!dbgL0) { int r = alloc(&dbgL0, v); if (r) return r; }
    if (!dbgL3) { int r = alloc(&dbgL3, v); if (r) return r; }
    if (!dbgLAll) { int r = alloc(&dbgLAll, v); if (r) return r; }
    return SS_OK;
  }
  void mark(const char *name) {
    if (!profiling) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    pev.push_back(e);
    pnames.push_back(name);
  }
};

namespace {

int upload(Core *e, float **dst, const float *src, size_t n) {
  int r = e->alloc(dst, n);
  if (r) return r;
  CK(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  return SS_OK;
}

int create_impl(Core *e, const float *mapLx, const float *mapLy, const float *mapRx,
                const float *mapRy, const float *a1, const float *a2, const float *a3) {
  const ss_config &c = e->cfg;
  std::vector<float> ha1, ha3; // matrix calibration: host copies of a1 / a3, only for the banded-output bound below
  if (e->has_cal && c.registration) {
    const size_t n = (size_t)c.rows * c.cols;
    ha1.resize(n); ha3.resize(n);
    const double *m = e->cal.reg_m;
    for (uint32_t v = 0; v < c.rows; ++v)
      for (uint32_t u = 0; u < c.cols; ++u) {
        ha1[(size_t)v * c.cols + u] = (float)((m[0] * u + m[1] * v) + m[2]);
        ha3[(size_t)v * c.cols + u] = (float)((m[6] * u + m[7] * v) + m[8]);
      }
    a1 = ha1.data(); a3 = ha3.data();
  }
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->cpy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->up, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->fr, cudaStreamNonBlocking));
  for (cudaEvent_t *ev : {&e->ev_upL, &e->ev_upR, &e->ev_lastband[0], &e->ev_lastband[1], &e->ev_rawfree[0], &e->ev_rawfree[1], &e->ev_done[0], &e->ev_done[1]})
    CK(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  for (auto &ev : e->ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : e->ev_seg) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : e->ev_band) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_cost, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_front, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_dep, cudaEventDisableTiming));
  const size_t fsz = e->fsz(), N = (size_t)c.batch;
  int r;
  if (!c.rectified && !e->has_cal) {
    if (!mapLx || !mapLy || !mapRx || !mapRy) return fail(SS_ERR_INVALID, "rectification maps required when rectified is false");
    if ((r = upload(e, &e->mapLx, mapLx, fsz))) return r;
    if ((r = upload(e, &e->mapLy, mapLy, fsz))) return r;
    if ((r = upload(e, &e->mapRx, mapRx, fsz))) return r;
    if ((r = upload(e, &e->mapRy, mapRy, fsz))) return r;
  }
  if (c.registration && !e->has_cal) {
    if (!a1 || !a2 || !a3) return fail(SS_ERR_INVALID, "registration planes a1,a2,a3 required");
    if ((r = upload(e, &e->a1, a1, fsz))) return r;
    if ((r = upload(e, &e->a2, a2, fsz))) return r;
    if ((r = upload(e, &e->a3, a3, fsz))) return r;
  }
  // Banded output needs, per matched column, the left-most RGB column it can reach (depth_and_splat in
  // post.cu: x = (a1 z + b1) / (a3 z + b3), z in [f b / D, inf) -- monotone in z while a3 z + b3 > 0).
  e->rgb_sufmin.clear();
  if (c.registration) {
    const double zmin = (double)c.focal_len * c.baseline_len / (double)c.max_disp;
    bool ok = zmin > 0;
    if (c.b3 > 0) { // invalid disparities give z = 0 and splat zRgb = b3 at (b1/b3, b2/b3) (camera.cu:187-195):
      // if that pixel is inside the RGB image any band may write it, so no RGB column is ever final early
      const double u0 = std::round((double)c.b1 / c.b3), v0 = std::round((double)c.b2 / c.b3);
      if (u0 >= 0 && u0 < (double)c.rgb_cols && v0 >= 0 && v0 < (double)c.rgb_rows) ok = false;
    }
    std::vector<int> colmin(c.cols, INT32_MAX);
    for (uint32_t y = 0; ok && y < c.rows; ++y)
      for (uint32_t x = 0; x < c.cols; ++x) {
        const size_t i = (size_t)y * c.cols + x;
        const double den0 = (double)a3[i] * zmin + c.b3;
        if (!(a3[i] > 0) || !(den0 > 0)) { ok = false; break; }
        const double x0 = ((double)a1[i] * zmin + c.b1) / den0, x1 = (double)a1[i] / (double)a3[i];
        const double lo = std::floor(std::min(x0, x1)) - 1.0;
        colmin[x] = std::min<double>(colmin[x], std::max(lo, -1.0e9));
      }
    if (ok) {
      e->rgb_sufmin.assign(c.cols + 1, INT32_MAX);
      for (int x = (int)c.cols - 1; x >= 0; --x) e->rgb_sufmin[x] = std::min(e->rgb_sufmin[x + 1], colmin[x]);
    }
  }
  // cost volumes: C, L1, L2 (S3 aliases L2) -- or 7 separate ones with keep_stages
  const size_t vol = fsz * (size_t)c.max_disp * sizeof(uint16_t);
  // (configurations outside the packed-u16 regime aggregate through generic.cu, which needs three more volumes)
  const bool fast0 = aggr_fast_supported(c.max_disp, census_bits(c.census_width, c.census_height) * c.bf_width * c.bf_height,
                                         c.p1 * c.bf_width * c.bf_height, c.p2 * c.bf_width * c.bf_height);
  const int nvol = c.keep_stages ? 7 : (fast0 ? 4 : 6); // C, L1, L2, S3 (S3 is its own volume: L1 / L2 may be 12-bit packed)
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t small = N * (fsz * 42 + e->rsz() * 52) + (64u << 20);
  if (free_b < small + (size_t)nvol * vol) return fail(SS_ERR_CUDA, "not enough device memory for one frame");
  size_t budget = (size_t)((double)(free_b - small) * 0.6);
  e->wave = (int)std::min<size_t>(N, std::max<size_t>(1, budget / ((size_t)nvol * vol)));
  if (g_max_wave > 0) e->wave = std::min(e->wave, g_max_wave); // (test hook: exercise the multi-wave path on a big GPU)
  const size_t wv = (size_t)e->wave * fsz * c.max_disp;
  if ((r = e->alloc(&e->C, wv))) return r;
  if ((r = e->alloc(&e->L1, wv))) return r;
  if ((r = e->alloc(&e->L2, wv))) return r;
  if ((r = e->alloc(&e->S3, wv))) return r;
  if (c.keep_stages && (r = e->ensure_generic_volumes())) return r;
  for (int k = 0; k < 2; ++k) {
    if ((r = e->alloc(&e->raw0s[k], N * fsz))) return r;
    if ((r = e->alloc(&e->raw1s[k], N * fsz))) return r;
    if ((r = e->alloc(&e->out_pair[k], N * e->rsz()))) return r;
  }
  e->raw0 = e->raw0s[0]; e->raw1 = e->raw1s[0]; e->out = e->out_pair[0];
  if ((r = e->alloc(&e->im0, N * fsz))) return r;
  if ((r = e->alloc(&e->im1, N * fsz))) return r;
  if ((r = e->alloc(&e->cen0, N * fsz))) return r;
  if ((r = e->alloc(&e->cen1, N * fsz))) return r;
  if ((r = e->alloc(&e->dispL, N * fsz))) return r;
  if ((r = e->alloc(&e->dispR, N * fsz))) return r;
  if ((r = e->alloc(&e->disp_lr, N * fsz))) return r;
  if ((r = e->alloc(&e->disp_med, N * fsz))) return r;
  if ((r = e->alloc(&e->disp_full, N * fsz))) return r;
  if ((r = e->alloc(&e->depth, N * fsz))) return r;
  if (c.registration && (r = e->alloc(&e->canvas_pair[0], N * e->rsz()))) return r;
  if (c.registration && (r = e->alloc(&e->canvas_pair[1], N * e->rsz()))) return r;
  e->canvas = e->canvas_pair[0];
  if ((r = e->alloc(&e->pc, N * e->rsz() * 3))) return r;
  if ((r = e->alloc(&e->rgbpc, N * e->rsz() * 6))) return r;
  CK(cudaStreamSynchronize(e->stream));
  return SS_OK;
}

enum InputKind { IN_U8, IN_RGBA };



// host_left / host_right: when non-null (host-u8 path, one wave, 7x7 census) the uploads happen here,
// the right image on the helper stream, so that the left image's front-end overlaps the second upload
// inputs_on_main: left/right are engine-owned buffers whose uploads were enqueued on the main stream
// async_out: asynchronous host frame (ss_submit_host_u8): this frame's depth map goes to *async_out (may be null for
// "no host delivery") and the main stream is NOT made to wait for the delivery -- ev_done[frame & 1] stands for it
int compute_impl(Core *e, InputKind kind, const void *left, const void *right,
                 const ss_bbox *bbox, cudaStream_t user, const uint8_t *host_left = nullptr,
                 const uint8_t *host_right = nullptr, bool inputs_on_main = false,
                 size_t env_pitch = 0, size_t row_pitch = 0, // pitches in source elements, 0 = packed
                 bool async = false, float *async_out = nullptr) {
  NvtxRange nvtx_frame("ss_b200::compute");
  const ss_config &c = e->cfg;
  if (!left || !right) return fail(SS_ERR_INVALID, "null input image");
  int bx = 0, by = 0, rows = (int)c.rows, cols = (int)c.cols, use_bbox = 0;
  if (bbox && bbox->enabled) {
    if (bbox->width < 1 || bbox->height < 1 || (uint64_t)bbox->x + bbox->width > c.cols ||
        (uint64_t)bbox->y + bbox->height > c.rows)
      return fail(SS_ERR_INVALID, "bbox must be non-empty and inside the image");
    if (c.bf_width / 2 >= (int)bbox->width || c.bf_height / 2 >= (int)bbox->height)
      return fail(SS_ERR_INVALID, "bbox smaller than the matching block");
    bx = (int)bbox->x; by = (int)bbox->y; cols = (int)bbox->width; rows = (int)bbox->height;
    use_bbox = 1;
  }
  e->computed = false;
  cudaStream_t st = e->stream;
  if (user && user != st) { // order after the caller's stream
    CK(cudaEventRecord(e->ev_in, user));
    CK(cudaStreamWaitEvent(st, e->ev_in, 0));
  }
  if (e->dep_pending) { // (the main stream already waits on it, see ss_wait_stream; the front-end may run on the helper stream)
    CK(cudaStreamWaitEvent(e->aux, e->ev_dep, 0));
    e->dep_pending = false;
  }
  if (e->pev.size() > 4096) { // profiling left on without anybody reading: drop the backlog
    for (auto ev : e->pev) cudaEventDestroy(ev);
    e->pev.clear(); e->pnames.clear();
  }
  e->mark("begin");

  const int slot = (int)(e->frame & 1);
  if (c.registration) e->canvas = e->canvas_pair[slot];
  e->out = e->out_pair[slot];
  // the previous frame's column bands run on the helper stream; they read the disparity maps that this frame's final
  // pass overwrites, so the main stream joins them right before that pass (and not earlier: the cost volume and the
  // three plain passes of this frame overlap them)
  auto join_prev_bands = [&]() -> int {
    if (e->band_pending) {
      CK(cudaStreamWaitEvent(st, e->ev_lastband[e->band_slot], 0));
      e->band_pending = false;
    }
    return SS_OK;
  };
  const int D = c.max_disp;
  const int P1 = c.p1 * c.bf_width * c.bf_height, P2 = c.p2 * c.bf_width * c.bf_height; // core.cu:670-671
  const int cmax = census_bits(c.census_width, c.census_height) * c.bf_width * c.bf_height;
  const bool fast = aggr_fast_supported(D, cmax, P1, P2);
  if (!fast) { int r = e->ensure_generic_volumes(); if (r) return r; }
  const int pack12 = fast && !c.keep_stages && g_allow_pack12 && aggr_pack12_supported(D, cmax, P2);
  e->last_pack12 = pack12 != 0;
  const size_t fsz = e->fsz(), msz = (size_t)rows * cols;
  int launches = 0;
  // Banded output plan (see ss_bind_output_host): up to 4 progress points of the final pass -> up to 5 bands
  int nseg = 0, seg_end[Core::MAXSEG];
  bool banded = false;
  // Host-input frames have a host consumer: without a bound buffer they stream into the engine's own pinned
  // staging buffer and ss_get_depth_host copies from there (a pageable 8.3 MB cudaMemcpy costs ~2 ms at C1).
  float *band_dst = async ? async_out : (e->host_out ? e->host_out : ((host_left || inputs_on_main) ? e->staging : nullptr));
  if (band_dst && fast && !use_bbox && !c.keep_stages && c.batch <= e->wave &&
      (!c.registration || !e->rgb_sufmin.empty())) {
    // progress points (3, 4 and 5 points measured within 1 % of each other on C1: the copy engine is the bound --
    // ~155 us for the map, starting when the first band is final -- not the band count)
    static const float frac[4] = {0.25f, 0.5f, 0.75f, 0.95f};
    for (float f : frac) {
      const int x = ((int)(cols * f) / 32) * 32;
      if (x - D - c.mf_size / 2 >= 32 && x > (nseg ? seg_end[nseg - 1] : 0) && x < cols) seg_end[nseg++] = x;
    }
    if (nseg > 0) {
      banded = true; // (the last band, up to cols, follows the end of the kernel)
      if (!e->progress) { int r = e->alloc(&e->progress, Core::MAXSEG); if (r) return r; }
    }
  }
  for (int w0 = 0; w0 < c.batch; w0 += e->wave) {
    const int wn = std::min(e->wave, c.batch - w0);
    FrontParams fp{};
    const size_t epx = kind == IN_U8 ? 1 : 4; // source elements per texel
    fp.src_env = env_pitch ? env_pitch : fsz * epx;
    fp.src_row = (uint32_t)(row_pitch ? row_pitch : (size_t)c.cols * epx);
    if (kind == IN_U8) {
      fp.left_u8 = static_cast<const uint8_t *>(left) + (size_t)w0 * fp.src_env;
      fp.right_u8 = static_cast<const uint8_t *>(right) + (size_t)w0 * fp.src_env;
    } else {
      fp.left_rgba = static_cast<const float *>(left) + (size_t)w0 * fp.src_env;
      fp.right_rgba = static_cast<const float *>(right) + (size_t)w0 * fp.src_env;
    }
    fp.mapLx = e->mapLx; fp.mapLy = e->mapLy; fp.mapRx = e->mapRx; fp.mapRy = e->mapRy;
    if (e->has_cal && !c.rectified) {
      fp.cal_maps = 1;
      std::memcpy(fp.rinvL, e->cal.rect_inv_left, sizeof(fp.rinvL));
      std::memcpy(fp.rinvR, e->cal.rect_inv_right, sizeof(fp.rinvR));
      fp.ir_fx = e->cal.ir_fx; fp.ir_fy = e->cal.ir_fy; fp.ir_cx = e->cal.ir_cx; fp.ir_cy = e->cal.ir_cy;
    }
    fp.frows = (int)c.rows; fp.fcols = (int)c.cols; fp.bx = bx; fp.by = by;
    fp.rows = rows; fp.cols = cols; fp.cw = c.census_width; fp.ch = c.census_height; fp.N = wn;
    fp.im0 = e->im0 + (size_t)w0 * msz; fp.im1 = e->im1 + (size_t)w0 * msz;
    fp.census0 = e->cen0 + (size_t)w0 * msz; fp.census1 = e->cen1 + (size_t)w0 * msz;
    fp.speckle_shape = c.speckle_shape; fp.speckle_scale = c.speckle_scale;
    fp.gaussian_mu = c.gaussian_mu; fp.gaussian_sigma = c.gaussian_sigma;
    fp.seed = c.ir_noise_seed; fp.frame = e->frame_id;
    if (c.registration && w0 == 0) { // the canvas of the whole batch is filled by the first wave's front-end
      fp.canvas = e->canvas; fp.canvas_n = (size_t)c.batch * e->rsz(); fp.canvas_fill = c.max_depth;
    }
    fp.only_image = -1;
    {
    NvtxRange nvtx_front("front");
    if (host_left) {
      // Host images: uploaded on the upload stream into raw slot `slot` (the front-end that last read it, two frames
      // ago, has finished), each image's front-end on the front stream as soon as its upload has landed and the
      // previous frame no longer needs the census buffers -- i.e. while the previous frame is still aggregating.
      const size_t bytes = (size_t)c.batch * fsz;
      fp.left_u8 = e->raw0s[slot]; fp.right_u8 = e->raw1s[slot];
      CK(cudaStreamWaitEvent(e->up, e->ev_rawfree[slot], 0));
      CK(cudaMemcpyAsync(e->raw0s[slot], host_left, bytes, cudaMemcpyHostToDevice, e->up));
      CK(cudaEventRecord(e->ev_upL, e->up));
      CK(cudaMemcpyAsync(e->raw1s[slot], host_right, bytes, cudaMemcpyHostToDevice, e->up));
      CK(cudaEventRecord(e->ev_upR, e->up));
      CK(cudaStreamWaitEvent(e->fr, e->ev_cost, 0)); // (first frame: never recorded, no-op)
      CK(cudaStreamWaitEvent(e->fr, e->ev_lastband[slot], 0)); // the dilation that last read this slot's splat canvas (two frames ago)
      CK(cudaStreamWaitEvent(e->fr, e->ev_upL, 0));
      fp.only_image = 0;
      CK(launch_front(fp, e->fr));
      CK(cudaStreamWaitEvent(e->fr, e->ev_upR, 0));
      fp.only_image = 1; fp.canvas = nullptr;
      CK(launch_front(fp, e->fr));
      CK(cudaEventRecord(e->ev_front, e->fr));
      CK(cudaEventRecord(e->ev_rawfree[slot], e->fr));
      CK(cudaStreamWaitEvent(st, e->ev_front, 0));
      ++launches;
    } else if (c.batch <= e->wave) {
      // Device inputs, one wave: the front-end runs on the helper stream and only waits until the previous
      // frame is past its aggregation, so with frames enqueued back to back it overlaps the previous
      // frame's post-processing (both are small latency-bound kernels: 24 us hidden, C1 1880 -> 1996
      // frames/s).  Released earlier it costs more than it hides: after the cost kernel -- the last reader
      // of the census buffers -- it takes bandwidth and the L2 lines the bottom->top pass counts on
      // (138 -> 166 us); after that pass it slows the final one (104 -> 123 us).  It fills the OTHER splat
      // canvas (the previous frame's dilation may still be reading its own).
      if (user && user != st) CK(cudaStreamWaitEvent(e->aux, e->ev_in, 0));
      if (inputs_on_main) { // the uploads into raw0/raw1 sit on the main stream: the helper stream must see them
        CK(cudaEventRecord(e->ev[0], st));
        CK(cudaStreamWaitEvent(e->aux, e->ev[0], 0));
      }
      CK(cudaStreamWaitEvent(e->aux, e->ev_cost, 0)); // (first frame: never recorded, no-op)
      CK(launch_front(fp, e->aux));
      CK(cudaEventRecord(e->ev_front, e->aux));
      CK(cudaStreamWaitEvent(st, e->ev_front, 0));
    } else {
      CK(launch_front(fp, st));
    }
    }
    e->mark("front");
    {
      NvtxRange nvtx_cost("cost");
      CK(launch_cost(fp.census0, fp.census1, e->C, wn, rows, cols, D, c.bf_width, c.bf_height,
                     census_bits(c.census_width, c.census_height), st));
    }
    e->mark("cost");
    NvtxRange nvtx_aggr("aggregation+wta");
    AggrBuffers ab{};
    ab.C = e->C; ab.L1 = e->L1; ab.L2 = e->L2; ab.S3 = e->S3;
    ab.dbgL0 = e->dbgL0; ab.dbgL3 = e->dbgL3; ab.dbgLAll = e->dbgLAll;
    ab.dispL = e->dispL + (size_t)w0 * msz; ab.dispR = e->dispR + (size_t)w0 * msz;
    ab.pack12 = pack12;
    if (fast) {
      if (!c.keep_stages) { ab.dbgL0 = ab.dbgL3 = ab.dbgLAll = nullptr; }
      AggrMarks am{[](void *ctx, const char *name) { static_cast<Core *>(ctx)->mark(name); }, e};
      CK(launch_aggr_passes(ab, wn, rows, cols, D, P1, P2, c.uniq_ratio, st, e->aux, e->ev, e->profiling ? &am : nullptr));
      launches += 3;
      if (!banded) { // (banded: the final pass follows below, with its progress counters)
        { int r = join_prev_bands(); if (r) return r; }
        CK(launch_aggr_final(ab, wn, rows, cols, D, P1, P2, c.uniq_ratio, st, nullptr, 0, nullptr));
        e->mark("aggr_right_wta");
        launches += 1;
      }
      // from here on the next frame's front-end may run (see above): what is left of this frame are the
      // small post-processing kernels (banded: the final pass as well)
      if (c.batch <= e->wave) CK(cudaEventRecord(e->ev_cost, st));
      launches += 2;
    } else {
      { int r = join_prev_bands(); if (r) return r; }
      CK(launch_aggr_wta_generic(ab, nullptr, wn, rows, cols, D, P1, P2, c.uniq_ratio, st));
      if (c.batch <= e->wave) CK(cudaEventRecord(e->ev_cost, st));
      launches += 8;
    }
    if (!fast) e->mark("aggr_wta_generic");
  }
  NvtxRange nvtx_post("post");
  PostParams pp{};
  pp.N = c.batch; pp.rows = rows; pp.cols = cols; pp.frows = (int)c.rows; pp.fcols = (int)c.cols;
  pp.bx = bx; pp.by = by; pp.bbox = use_bbox; pp.lr_max_diff = c.lr_max_diff; pp.mf_size = c.mf_size;
  pp.focal = c.focal_len; pp.baseline = c.baseline_len; pp.min_depth = c.min_depth; pp.max_depth = c.max_depth;
  pp.dispL = e->dispL; pp.dispR = e->dispR;
  pp.disp_lr = c.keep_stages ? e->disp_lr : nullptr;
  pp.disp_med = e->disp_med;
  pp.disp_full = use_bbox ? e->disp_full : e->disp_med;
  pp.depth = e->depth;
  pp.registration = c.registration; pp.dilation = c.dilation;
  pp.a1 = e->a1; pp.a2 = e->a2; pp.a3 = e->a3; pp.b1 = c.b1; pp.b2 = c.b2; pp.b3 = c.b3;
  if (e->has_cal) std::memcpy(pp.reg_m, e->cal.reg_m, sizeof(pp.reg_m));
  pp.rgb_rows = (int)c.rgb_rows; pp.rgb_cols = (int)c.rgb_cols;
  pp.canvas = e->canvas; pp.out = e->out;
  pp.canvas_prefilled = 1;
  e->streamed = false;
  e->stream_dst = nullptr;
  if (banded) {
    // final pass in column segments; behind each segment: post-processing of the columns whose
    // disparities are final (helper stream) and the copy of the finished output columns to the bound
    // host buffer (copy stream), both overlapping the next segment
    AggrBuffers ab{};
    ab.C = e->C; ab.L1 = e->L1; ab.L2 = e->L2; ab.S3 = e->S3;
    ab.dispL = e->dispL; ab.dispR = e->dispR;
    ab.pack12 = pack12;
    const int H = c.mf_size / 2, ocols = (int)e->out_cols(), orows = (int)e->out_rows();
    const size_t pitch = (size_t)ocols * sizeof(float);
    { int r = join_prev_bands(); if (r) return r; } // (also: nobody polls the progress counters any more)
    CK(cudaStreamWaitEvent(e->aux, e->ev_done[slot], 0)); // the read-back of the frame that last used this output slot
    CK(cudaMemsetAsync(e->progress, 0, Core::MAXSEG * sizeof(uint32_t), st));
    CK(cudaEventRecord(e->ev_seg[0], st));
    CK(cudaStreamWaitEvent(e->aux, e->ev_seg[0], 0)); // counters are zero before anybody waits on them
    CK(launch_aggr_final(ab, c.batch, rows, cols, D, P1, P2, c.uniq_ratio, st, e->progress, nseg, seg_end));
    ++launches;
    CK(cudaEventRecord(e->ev_seg[1], st)); // end of the pass: the last band
    int xa = 0, ua = 0;
    for (int s = 0; s <= nseg; ++s) {
      const bool last = s == nseg;
      if (last) { // the columns behind the last progress point: after the pass itself
        CK(cudaStreamWaitEvent(e->aux, e->ev_seg[1], 0));
      } else { // a one-warp gate holds the helper stream until every row has passed seg_end[s]
        CK(launch_gate(e->progress + s, (uint32_t)c.batch * (uint32_t)rows, e->aux));
        ++launches;
      }
      // LR check reads dispR up to D-1 columns back (final D-1 columns behind the pass), the median H columns ahead
      const int xb = last ? cols : std::max(xa, ((seg_end[s] - D - H) / 32) * 32);
      int ub;
      if (last) ub = ocols;
      else if (c.registration) ub = std::min(ocols, std::max(ua, ((e->rgb_sufmin[xb] - 2) / 4) * 4)); // dilation reads one column ahead
      else ub = xb;
      pp.xa = xa; pp.xb = xb; pp.ua = c.registration ? ua : 0; pp.ub = c.registration ? ub : 0;
      if (xb > xa || (c.registration && ub > ua)) {
        int pl = 0;
        CK(launch_post(pp, e->aux, &pl));
        launches += pl;
      }
      CK(cudaEventRecord(e->ev_band[s], e->aux));
      if (ub > ua) {
        CK(cudaStreamWaitEvent(e->cpy, e->ev_band[s], 0));
        CK(cudaMemcpy2DAsync(band_dst + ua, pitch, e->out + ua, pitch, (size_t)(ub - ua) * sizeof(float),
                             (size_t)c.batch * orows, cudaMemcpyDeviceToHost, e->cpy));
      }
      xa = xb; ua = ub;
    }
    e->mark("aggr_right_wta");
    CK(cudaEventRecord(e->ev_lastband[slot], e->aux));
    CK(cudaEventRecord(e->ev_done[slot], e->cpy));
    e->band_pending = true;
    e->band_slot = slot;
    if (!async) { // synchronous callers: completion of the main stream means "delivered"
      int r = join_prev_bands(); if (r) return r;
      CK(cudaStreamWaitEvent(st, e->ev_done[slot], 0));
    }
    e->streamed = true;
    e->stream_dst = band_dst;
  } else {
    int pl = 0;
    CK(launch_post(pp, st, &pl));
    launches += pl;
    if (async) { // no column bands for this configuration: one copy behind the frame
      if (async_out) CK(cudaMemcpyAsync(async_out, e->out, (size_t)c.batch * e->rsz() * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(e->ev_done[slot], st));
      e->streamed = async_out != nullptr;
      e->stream_dst = async_out;
    }
  }
  if (async) e->slot_ticket[slot] = e->frame + 1;
  e->async_unwaited = async;
  e->mark("post");
  e->launches = launches;
  e->mrows = rows; e->mcols = cols;
  e->frame++;
  if (user && user != st) {
    CK(cudaEventRecord(e->ev_out, st));
    CK(cudaStreamWaitEvent(user, e->ev_out, 0));
  }
  e->computed = true;
  return SS_OK;
}

} // namespace

extern "C" int ssb_debug_set_pack12(int on) { g_allow_pack12 = on; return 0; }
extern "C" int ssb_debug_set_max_wave(int n) { g_max_wave = n; return 0; }

// Host-facing calls that read results through the main stream first join whatever an asynchronous frame left on the
// helper and copy streams (its column bands and their copies).
static int join_async(Core *e) {
  if (e->band_pending) {
    CK(cudaStreamWaitEvent(e->stream, e->ev_lastband[e->band_slot], 0));
    CK(cudaStreamWaitEvent(e->stream, e->ev_done[e->band_slot], 0));
    e->band_pending = false;
  }
  return SS_OK;
}

// ---- one lane: the C-ABI operations on a single Core (the public entry points are at the end of the file) ----

static int create_common(const ss_config *cfg, const ss_calibration *cal, const float *mapLx, const float *mapLy,
                         const float *mapRx, const float *mapRy, const float *a1, const float *a2, const float *a3,
                         Core **out);
static int core_destroy(Core *e);

static int core_create(const ss_config *cfg, const float *mapLx, const float *mapLy, const float *mapRx,
              const float *mapRy, const float *a1, const float *a2, const float *a3,
              Core **out) {
  return create_common(cfg, nullptr, mapLx, mapLy, mapRx, mapRy, a1, a2, a3, out);
}

static int core_create_calibrated(const ss_config *cfg, const ss_calibration *cal, Core **out) {
  if (!cal) return fail(SS_ERR_INVALID, "null calibration");
  for (double v : cal->reg_m) if (!std::isfinite(v)) return fail(SS_ERR_INVALID, "registration matrix is not finite");
  if (cfg && !cfg->rectified) {
    for (int i = 0; i < 9; ++i)
      if (!std::isfinite(cal->rect_inv_left[i]) || !std::isfinite(cal->rect_inv_right[i])) return fail(SS_ERR_INVALID, "rectification matrix is not finite");
    if (!(cal->ir_fx > 0) || !(cal->ir_fy > 0)) return fail(SS_ERR_INVALID, "IR focal lengths must be positive");
  }
  return create_common(cfg, cal, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out);
}

static int create_common(const ss_config *cfg, const ss_calibration *cal, const float *mapLx, const float *mapLy,
                         const float *mapRx, const float *mapRy, const float *a1, const float *a2, const float *a3,
                         Core **out) {
  if (!cfg || !out) return fail(SS_ERR_INVALID, "null argument");
  *out = nullptr;
  int r = validate_params(*cfg);
  if (r) return r;
  if (cfg->batch < 1) return fail(SS_ERR_INVALID, "batch must be >= 1");
  if (cfg->registration && (cfg->rgb_rows < 1 || cfg->rgb_cols < 1)) return fail(SS_ERR_INVALID, "RGB resolution must be positive");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(SS_ERR_NO_DEVICE, "no CUDA device: the B200 stereo depth engine has no CPU fallback");
  int dev = cfg->device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= count) return fail(SS_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(SS_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100: this build targets B200 only");
  DeviceGuard g(dev);
  if (!g.ok) return fail(SS_ERR_CUDA, "cudaSetDevice failed");
  Core *e = new Core();
  e->cfg = *cfg;
  if (e->cfg.lr_max_diff == -1) e->cfg.lr_max_diff = 255; // uint8_t wrap in the reference ctor
  e->device = dev;
  e->cfg.device = dev;
  if (cal) { e->has_cal = true; e->cal = *cal; }
  r = create_impl(e, mapLx, mapLy, mapRx, mapRy, a1, a2, a3);
  if (r) { std::string keep = g_err; core_destroy(e); g_err = keep; return r; }
  *out = e;
  return SS_OK;
}

static int core_destroy(Core *e) {
  if (!e) return SS_OK;
  DeviceGuard g(e->device);
  for (cudaStream_t s : {e->stream, e->aux, e->cpy, e->up, e->fr})
    if (s) cudaStreamSynchronize(s);
  for (void *p : e->allocs) cudaFree(p);
  if (e->staging) cudaFreeHost(e->staging);
  for (auto ev : e->pev) cudaEventDestroy(ev);
  for (auto ev : e->ev) if (ev) cudaEventDestroy(ev);
  for (auto ev : e->ev_seg) if (ev) cudaEventDestroy(ev);
  for (auto ev : e->ev_band) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : {e->ev_upL, e->ev_upR, e->ev_lastband[0], e->ev_lastband[1], e->ev_rawfree[0], e->ev_rawfree[1], e->ev_done[0], e->ev_done[1]})
    if (ev) cudaEventDestroy(ev);
  if (e->up) { cudaStreamSynchronize(e->up); cudaStreamDestroy(e->up); }
  if (e->fr) { cudaStreamSynchronize(e->fr); cudaStreamDestroy(e->fr); }
  if (e->ev_cost) cudaEventDestroy(e->ev_cost);
  if (e->ev_front) cudaEventDestroy(e->ev_front);
  if (e->cpy) { cudaStreamSynchronize(e->cpy); cudaStreamDestroy(e->cpy); }
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->ev_dep) cudaEventDestroy(e->ev_dep);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->aux) cudaStreamDestroy(e->aux);
  delete e;
  return SS_OK;
}

static int core_compute_host_u8(Core *e, const uint8_t *left, const uint8_t *right, const ss_bbox *bbox) {
  if (!e || !left || !right) return fail(SS_ERR_INVALID, "null argument");
  DeviceGuard g(e->device);
  const size_t bytes = (size_t)e->cfg.batch * e->fsz();
  const bool split = e->cfg.census_width == 7 && e->cfg.census_height == 7 && e->cfg.batch <= e->wave;
  if (!e->host_out && !e->staging) { // first host-input frame without a bound output: allocate the pinned staging buffer
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)e->cfg.batch * e->rsz() * sizeof(float), cudaHostAllocDefault) == cudaSuccess) e->staging = static_cast<float *>(p);
    else cudaGetLastError(); // (no pinned memory to spare: the read-back falls back to a direct copy)
  }
  if (!split) {
    CK(cudaMemcpyAsync(e->raw0, left, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->raw1, right, bytes, cudaMemcpyHostToDevice, e->stream));
  }
  int r = compute_impl(e, IN_U8, e->raw0, e->raw1, bbox, nullptr, split ? left : nullptr, split ? right : nullptr, !split);
  if (r) return r;
  CK(cudaStreamSynchronize(e->stream));
  return SS_OK;
}

static int core_compute_device_rgba_f32(Core *e, const void *left, const void *right, const ss_bbox *bbox, void *stream) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  DeviceGuard g(e->device);
  return compute_impl(e, IN_RGBA, left, right, bbox, static_cast<cudaStream_t>(stream));
}

static int core_compute_device_rgba_f32_pitched(Core *e, const void *left, const void *right, size_t env_pitch_bytes,
                                       size_t row_pitch_bytes, const ss_bbox *bbox, void *stream) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  const size_t row_min = (size_t)e->cfg.cols * 16;
  if (row_pitch_bytes < row_min || row_pitch_bytes % 4 || row_pitch_bytes / 4 > 0xffffffffull ||
      (e->cfg.rows - 1) * (row_pitch_bytes / 4) + (size_t)e->cfg.cols * 4 > 0xffffffffull)
    return fail(SS_ERR_INVALID, "row pitch must be a multiple of 4 bytes, at least cols*16, and one image must span < 16 GiB");
  if (e->cfg.batch > 1 && (env_pitch_bytes % 4 || env_pitch_bytes < (e->cfg.rows - 1) * row_pitch_bytes + row_min))
    return fail(SS_ERR_INVALID, "environment pitch must be a multiple of 4 bytes and hold one image");
  if ((reinterpret_cast<uintptr_t>(left) | reinterpret_cast<uintptr_t>(right)) % 4)
    return fail(SS_ERR_INVALID, "float32 inputs must be 4-byte aligned");
  DeviceGuard g(e->device);
  return compute_impl(e, IN_RGBA, left, right, bbox, static_cast<cudaStream_t>(stream), nullptr, nullptr, false,
                      env_pitch_bytes / 4, row_pitch_bytes / 4);
}

static int core_compute_device_u8(Core *e, const void *left, const void *right, const ss_bbox *bbox, void *stream) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  DeviceGuard g(e->device);
  return compute_impl(e, IN_U8, left, right, bbox, static_cast<cudaStream_t>(stream));
}

static int core_wait_stream(Core *e, void *stream) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!s || s == e->stream) return SS_OK;
  DeviceGuard g(e->device);
  if (e->dep_pending) CK(cudaStreamWaitEvent(e->aux, e->ev_dep, 0)); // a second producer stream: flush the first
  CK(cudaEventRecord(e->ev_dep, s));
  CK(cudaStreamWaitEvent(e->stream, e->ev_dep, 0));
  e->dep_pending = true;
  return SS_OK;
}

static int core_submit_host_u8(Core *e, const uint8_t *left, const uint8_t *right, const ss_bbox *bbox, float *out_host,
                      size_t capacity_bytes, uint64_t *ticket) {
  if (!e || !left || !right || !ticket) return fail(SS_ERR_INVALID, "null argument");
  if (out_host && capacity_bytes < (size_t)e->cfg.batch * e->rsz() * sizeof(float)) return fail(SS_ERR_INVALID, "output buffer too small");
  DeviceGuard g(e->device);
  const int slot = (int)(e->frame & 1);
  if (e->slot_ticket[slot]) { // at most two frames in flight: the frame that last used this slot must have been delivered
    CK(cudaEventSynchronize(e->ev_done[slot]));
    e->slot_ticket[slot] = 0;
  }
  const size_t bytes = (size_t)e->cfg.batch * e->fsz();
  const bool split = e->cfg.census_width == 7 && e->cfg.census_height == 7 && e->cfg.batch <= e->wave;
  if (!split) { // generic census / multi-wave batches: uploads and frame on the main stream (asynchronous, not overlapped)
    CK(cudaMemcpyAsync(e->raw0, left, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->raw1, right, bytes, cudaMemcpyHostToDevice, e->stream));
  }
  int r = compute_impl(e, IN_U8, e->raw0, e->raw1, bbox, nullptr, split ? left : nullptr, split ? right : nullptr, !split, 0, 0,
                       true, out_host);
  if (r) return r;
  *ticket = e->frame; // (compute_impl has advanced the frame counter: ticket = frame number + 1)
  return SS_OK;
}

static int core_wait_frame(Core *e, uint64_t ticket) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (ticket == 0 || ticket > e->frame) return fail(SS_ERR_INVALID, "unknown frame ticket");
  DeviceGuard g(e->device);
  const int slot = (int)((ticket - 1) & 1);
  if (e->slot_ticket[slot] == ticket) { // (an older ticket of this slot was already waited for by a later submit)
    CK(cudaEventSynchronize(e->ev_done[slot]));
    e->slot_ticket[slot] = 0;
  }
  if (ticket == e->frame) e->async_unwaited = false;
  return SS_OK;
}

static int core_synchronize(Core *e) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  DeviceGuard g(e->device);
  { int r = join_async(e); if (r) return r; }
  CK(cudaStreamSynchronize(e->stream));
  e->async_unwaited = false;
  return SS_OK;
}

static int core_get_output_shape(const Core *e, uint32_t *rows, uint32_t *cols) {
  if (!e || !rows || !cols) return fail(SS_ERR_INVALID, "null argument");
  *rows = e->out_rows(); *cols = e->out_cols();
  return SS_OK;
}
static int core_get_input_shape(const Core *e, uint32_t *rows, uint32_t *cols) {
  if (!e || !rows || !cols) return fail(SS_ERR_INVALID, "null argument");
  *rows = e->cfg.rows; *cols = e->cfg.cols;
  return SS_OK;
}
static int core_get_stream(const Core *e, void **stream) {
  if (!e || !stream) return fail(SS_ERR_INVALID, "null argument");
  *stream = e->stream;
  return SS_OK;
}
static int core_get_device(const Core *e, int32_t *device) {
  if (!e || !device) return fail(SS_ERR_INVALID, "null argument");
  *device = e->device;
  return SS_OK;
}

static int copy_out(Core *e, const float *src, size_t count, float *out, size_t cap) {
  if (!out) return fail(SS_ERR_INVALID, "null output");
  if (cap < count * sizeof(float)) return fail(SS_ERR_INVALID, "output buffer too small");
  CK(cudaMemcpyAsync(out, src, count * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return SS_OK;
}

static int core_get_depth_host(Core *e, float *out, size_t cap) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  { int r = join_async(e); if (r) return r; }
  const size_t bytes = (size_t)e->cfg.batch * e->rsz() * sizeof(float);
  if (e->streamed && out && out == e->stream_dst) { // already delivered band by band during compute
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
  }
  if (e->streamed && out && e->stream_dst == e->staging && e->staging) { // delivered into the pinned staging buffer
    if (cap < bytes) return fail(SS_ERR_INVALID, "output buffer too small");
    CK(cudaStreamSynchronize(e->stream));
    std::memcpy(out, e->staging, bytes);
    return SS_OK;
  }
  return copy_out(e, e->out, (size_t)e->cfg.batch * e->rsz(), out, cap);
}

static int core_bind_output_host(Core *e, float *out, size_t cap) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (out && cap < (size_t)e->cfg.batch * e->rsz() * sizeof(float)) return fail(SS_ERR_INVALID, "output buffer too small");
  DeviceGuard g(e->device);
  { int r = join_async(e); if (r) return r; }
  CK(cudaStreamSynchronize(e->stream)); // a frame may still be streaming into the previous binding
  e->host_out = out;
  e->host_cap = out ? cap : 0;
  e->streamed = false;
  e->stream_dst = nullptr;
  return SS_OK;
}
static int core_get_depth_device(Core *e, void **ptr) {
  if (!e || !ptr) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  if (e->band_pending || e->async_unwaited) { // an asynchronous host frame nobody waited for: the getter waits (its caller
    DeviceGuard g(e->device);                  // has no stream the result could be ordered on)
    int r = join_async(e); if (r) return r;
    CK(cudaStreamSynchronize(e->stream));
    e->async_unwaited = false;
  }
  *ptr = e->out;
  return SS_OK;
}
static int run_pc(Core *e, const void *rgba) {
  const ss_config &c = e->cfg;
  { int r = join_async(e); if (r) return r; }
  if (rgba) { // the colour image may still be in flight on the caller's (default) stream
    CK(cudaEventRecord(e->ev_in, cudaStreamLegacy));
    CK(cudaStreamWaitEvent(e->stream, e->ev_in, 0));
  }
  CK(launch_point_cloud(e->out, static_cast<const float *>(rgba), rgba ? e->rgbpc : e->pc, c.batch,
                        (int)e->out_rows(), (int)e->out_cols(), c.main_fx, c.main_fy, c.main_skew,
                        c.main_cx, c.main_cy, e->stream));
  return SS_OK;
}
static int core_get_point_cloud_host(Core *e, float *out, size_t cap) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  int r = run_pc(e, nullptr);
  if (r) return r;
  return copy_out(e, e->pc, (size_t)e->cfg.batch * e->rsz() * 3, out, cap);
}
static int core_get_point_cloud_device(Core *e, void **ptr) {
  if (!e || !ptr) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  int r = run_pc(e, nullptr);
  if (r) return r;
  CK(cudaStreamSynchronize(e->stream)); // the reference's getter is synchronous (core.cu:409-411)
  *ptr = e->pc;
  return SS_OK;
}
// stream-ordered variants: the kernel is enqueued behind the frame, nothing waits on the host
static int core_enqueue_point_cloud(Core *e, const void *rgba, void **ptr) {
  if (!e || !ptr) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  int r = run_pc(e, rgba);
  if (r) return r;
  *ptr = rgba ? e->rgbpc : e->pc;
  return SS_OK;
}
static int core_get_rgb_point_cloud_host(Core *e, const void *rgba, float *out, size_t cap) {
  if (!e || !rgba) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  int r = run_pc(e, rgba);
  if (r) return r;
  return copy_out(e, e->rgbpc, (size_t)e->cfg.batch * e->rsz() * 6, out, cap);
}
static int core_get_rgb_point_cloud_device(Core *e, const void *rgba, void **ptr) {
  if (!e || !rgba || !ptr) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  DeviceGuard g(e->device);
  int r = run_pc(e, rgba);
  if (r) return r;
  CK(cudaStreamSynchronize(e->stream));
  *ptr = e->rgbpc;
  return SS_OK;
}

static int core_set_ir_noise_parameters(Core *e, float shape, float scale, float mu, float sigma) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  e->cfg.speckle_shape = shape; e->cfg.speckle_scale = scale;
  e->cfg.gaussian_mu = mu; e->cfg.gaussian_sigma = sigma;
  return SS_OK;
}
#define SET_VALIDATED(body)                                                                       \
  if (!e) return fail(SS_ERR_INVALID, "null engine");                                             \
  ss_config t = e->cfg;                                                                           \
  body;                                                                                           \
  int r = validate_params(t);                                                                     \
  if (r) return r;                                                                                \
  e->cfg = t;                                                                                     \
  return SS_OK;
static int core_set_penalties(Core *e, int32_t p1, int32_t p2) { SET_VALIDATED(t.p1 = p1; t.p2 = p2) }
static int core_set_census_window_size(Core *e, int32_t w, int32_t h) { SET_VALIDATED(t.census_width = w; t.census_height = h) }
static int core_set_matching_block_size(Core *e, int32_t w, int32_t h) { SET_VALIDATED(t.bf_width = w; t.bf_height = h) }
static int core_set_uniqueness_ratio(Core *e, int32_t u) { SET_VALIDATED(t.uniq_ratio = u) }
static int core_set_lr_max_diff(Core *e, int32_t d) { SET_VALIDATED(t.lr_max_diff = (d == -1 ? 255 : d)) }

static int core_get_stage_host(Core *e, const char *name, int32_t index, void *out, size_t cap, size_t *bytes) {
  if (!e || !name || !out) return fail(SS_ERR_INVALID, "null argument");
  if (!e->computed) return fail(SS_ERR_NOT_COMPUTED, "No computed data stored");
  if (index < 0 || index >= e->cfg.batch) return fail(SS_ERR_INVALID, "batch index out of range");
  DeviceGuard g(e->device);
  { int r = join_async(e); if (r) return r; }
  const std::string n(name);
  const size_t msz = (size_t)e->mrows * e->mcols, fsz = e->fsz(), D = (size_t)e->cfg.max_disp;
  const void *src = nullptr;
  size_t sz = 0;
  auto pix = [&](const void *base, size_t elem, size_t per) { src = static_cast<const char *>(base) + (size_t)index * per * elem; sz = per * elem; };
  auto vol = [&](const uint16_t *base) -> int {
    if (!base) return fail(SS_ERR_INVALID, "stage '" + n + "' needs keep_stages=1");
    if (e->cfg.batch > e->wave) return fail(SS_ERR_INVALID, "volume stages are only kept when the batch fits one wave");
    src = base + (size_t)index * msz * D; sz = msz * D * 2;
    return SS_OK;
  };
  int r = SS_OK;
  if (n == "im0") pix(e->im0, 1, msz);
  else if (n == "im1") pix(e->im1, 1, msz);
  else if (n == "census0") pix(e->cen0, 4, msz);
  else if (n == "census1") pix(e->cen1, 4, msz);
  else if (n == "cost") r = vol(e->C);
  else if (n == "L1") r = vol(e->L1);
  else if (n == "L2") r = vol(e->cfg.keep_stages ? e->L2 : nullptr);
  else if (n == "L0") r = vol(e->cfg.keep_stages ? e->dbgL0 : nullptr);
  else if (n == "L3") r = vol(e->cfg.keep_stages ? e->dbgL3 : nullptr);
  else if (n == "LAll") r = vol(e->cfg.keep_stages ? e->dbgLAll : nullptr);
  else if (n == "disp_wta") pix(e->dispL, 4, msz);
  else if (n == "disp_right") pix(e->dispR, 2, msz);
  else if (n == "disp_lr") { if (!e->cfg.keep_stages) return fail(SS_ERR_INVALID, "stage 'disp_lr' needs keep_stages=1"); pix(e->disp_lr, 4, msz); }
  else if (n == "disp_med") pix(e->disp_med, 4, msz);
  else if (n == "disp_full") pix(e->mrows == (int)e->cfg.rows && e->mcols == (int)e->cfg.cols ? e->disp_med : e->disp_full, 4, fsz);
  else if (n == "depth") pix(e->depth, 4, fsz);
  else return fail(SS_ERR_INVALID, "unknown stage '" + n + "'");
  if (r) return r;
  if (bytes) *bytes = sz;
  if (cap < sz) return fail(SS_ERR_INVALID, "output buffer too small");
  if (n == "L1" && e->last_pack12) { // 12-bit packed on the device (1.5 D bytes per pixel): unpack on the host
    const size_t pb = D * 3 / 2;
    std::vector<unsigned char> tmp(msz * pb);
    CK(cudaMemcpyAsync(tmp.data(), reinterpret_cast<const unsigned char *>(e->L1) + (size_t)index * msz * pb, tmp.size(),
                       cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    uint16_t *o = static_cast<uint16_t *>(out);
    for (size_t p = 0; p < msz; ++p)
      for (size_t d = 0; d < D; ++d) {
        const unsigned nib = (tmp[p * pb + D + d / 2] >> (4 * (d & 1))) & 0xfu;
        o[p * D + d] = (uint16_t)(tmp[p * pb + d] | (nib << 8));
      }
    return SS_OK;
  }
  CK(cudaMemcpyAsync(out, src, sz, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return SS_OK;
}

static int core_set_profiling(Core *e, int32_t enabled) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  e->profiling = enabled != 0;
  return SS_OK;
}
static int core_get_stage_times(Core *e, const char **names, float *ms, int32_t capacity, int32_t *count,
                       int32_t *frames) {
  if (!e || !count) return fail(SS_ERR_INVALID, "null argument");
  DeviceGuard g(e->device);
  CK(cudaStreamSynchronize(e->stream));
  std::vector<const char *> uniq;
  std::vector<float> tot;
  int nframes = 0;
  for (size_t i = 0; i < e->pev.size(); ++i) {
    if (std::strcmp(e->pnames[i], "begin") == 0) { ++nframes; continue; }
    float t = 0;
    CK(cudaEventElapsedTime(&t, e->pev[i - 1], e->pev[i]));
    size_t k = 0;
    while (k < uniq.size() && std::strcmp(uniq[k], e->pnames[i]) != 0) ++k;
    if (k == uniq.size()) { uniq.push_back(e->pnames[i]); tot.push_back(0.f); }
    tot[k] += t;
  }
  *count = (int32_t)uniq.size();
  if (frames) *frames = nframes;
  for (size_t k = 0; k < uniq.size() && (int32_t)k < capacity; ++k) {
    if (names) names[k] = uniq[k];
    if (ms) ms[k] = tot[k];
  }
  for (auto ev : e->pev) cudaEventDestroy(ev);
  e->pev.clear(); e->pnames.clear();
  return SS_OK;
}
static int core_get_launches_per_compute(Core *e, int32_t *count) {
  if (!e || !count) return fail(SS_ERR_INVALID, "null argument");
  *count = e->launches;
  return SS_OK;
}

// =====================================================================================================================
// The public engine: up to three LANES behind the C ABI.
//
// A lane (Core) owns a complete set of streams and buffers.  Single frames of one stereo pair leave every kernel with
// a latency-bound head and tail (a 720-row pass cannot fill 148 SMs evenly; the final pass is a serial chain per row),
// and kernels of one stream run back to back, so with ONE lane those phases add up.  With several lanes consecutive
// frames rotate over independent stream sets and the tail of one frame's kernel is filled by the other frames'
// kernels: C1 1 954 (one lane) -> 2 250 (two) -> 2 285 (three) frames/s with device inputs (tools/hook_ab.py
// ssb_debug_set_auto_lanes, round 2), bit-identical results; a lane costs four cost volumes of memory (1 GB at C1).
// Batched engines (many environments per call) already fill the machine and gain nothing: they get one lane.
//
// The engine's PUBLIC stream (ss_get_stream) runs no kernels: it waits, in submission order, for the end of every
// frame, so work enqueued on it after compute() sees that frame (and all earlier ones).  A frame starts only after the
// work that was on the public stream when the PREVIOUS frame was enqueued -- i.e. after every consumer of the frame that
// last used its lane -- which keeps "borrowed result pointer + stream order" safe without serialising the lanes.
// =====================================================================================================================
struct ss_engine {
  static constexpr int MAXL = 3;
  Core *lane[MAXL] = {nullptr, nullptr, nullptr};
  int nlanes = 1;
  int last = 0;          // lane of the most recent frame
  uint64_t frames = 0;   // frames enqueued so far
  int device = 0;
  bool profiling = false; // stage times are measured with the frames of ONE lane running alone
  cudaStream_t pub = nullptr;
  cudaEvent_t ev_join = nullptr, ev_tail[MAXL] = {nullptr, nullptr, nullptr};
  struct Ticket { uint64_t pub = 0, sub = 0; int lane = 0; } tk[8];

  int next_lane() const { return (profiling || nlanes == 1) ? 0 : (int)(frames % (uint64_t)nlanes); }
};

namespace {

int g_auto_lanes = 3; // lanes of a single-pair engine created with lanes = 0 (ssb_debug_set_auto_lanes: A/B hook, tools/hook_ab.py)

// before a frame is enqueued on `lane`: order it after the public stream as it was when the previous frame was enqueued
int lane_begin(ss_engine *e, int lane) {
  // (frame k reuses the lane of frame k - nlanes: it waits for the public stream as it was when frame k - nlanes + 1 was enqueued)
  const uint64_t back = (uint64_t)std::max(e->nlanes - 1, 1);
  if (e->frames >= back) CK(cudaStreamWaitEvent(e->lane[lane]->stream, e->ev_tail[(e->frames - back) % ss_engine::MAXL], 0));
  return SS_OK;
}
// after a frame has been enqueued on `lane`: the public stream completes behind it
int lane_end(ss_engine *e, int lane) {
  CK(cudaEventRecord(e->ev_tail[e->frames % ss_engine::MAXL], e->pub)); // (the public stream BEFORE this frame's join)
  CK(cudaEventRecord(e->ev_join, e->lane[lane]->stream));
  CK(cudaStreamWaitEvent(e->pub, e->ev_join, 0));
  e->last = lane;
  e->frames++;
  return SS_OK;
}
// caller streams: the public stream itself (or null) means "the inputs are complete, no ordering"
void *user_stream(const ss_engine *e, void *stream) { return stream == (void *)e->pub ? nullptr : stream; }

int create_lanes(const ss_config *cfg, const ss_calibration *cal, const float *mapLx, const float *mapLy, const float *mapRx,
                 const float *mapRy, const float *a1, const float *a2, const float *a3, ss_engine **out) {
  if (!cfg || !out) return fail(SS_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->lanes < 0 || cfg->lanes > ss_engine::MAXL) return fail(SS_ERR_INVALID, "lanes must be 0 (automatic), 1, 2 or 3");
  ss_engine *e = new ss_engine();
  auto make = [&](Core **c) {
    return cal ? core_create_calibrated(cfg, cal, c) : core_create(cfg, mapLx, mapLy, mapRx, mapRy, a1, a2, a3, c);
  };
  int r = make(&e->lane[0]);
  if (r) { delete e; return r; }
  e->device = e->lane[0]->device;
  const int want = cfg->lanes ? cfg->lanes : ((cfg->batch == 1 && !cfg->keep_stages) ? g_auto_lanes : 1);
  if (want >= 2) {
    ss_config c2 = *cfg;
    c2.device = e->device;
    const ss_config *saved = cfg;
    cfg = &c2;
    const std::string keep = g_err;
    for (int i = 1; i < want; ++i) {
      if (make(&e->lane[i]) == SS_OK) e->nlanes = i + 1; // (no memory for another lane: fewer lanes, silently)
      else { e->lane[i] = nullptr; g_err = keep; cudaGetLastError(); break; }
    }
    cfg = saved;
  }
  DeviceGuard g(e->device);
  cudaError_t ce = cudaStreamCreateWithFlags(&e->pub, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
  for (auto &ev : e->ev_tail) if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (ce != cudaSuccess) { ss_destroy(e); return fail(SS_ERR_CUDA, cudaGetErrorString(ce)); }
  *out = e;
  return SS_OK;
}

} // namespace

extern "C" {

void ssb_debug_set_auto_lanes(int n) { g_auto_lanes = n < 1 ? 1 : (n > ss_engine::MAXL ? ss_engine::MAXL : n); }
const char *ss_last_error(void) { return g_err.c_str(); }
const char *ss_version(void) { return "ss_b200 0.2 (sm_100a)"; }

int ss_alloc_host(size_t bytes, void **ptr) {
  if (!ptr || !bytes) return fail(SS_ERR_INVALID, "null argument");
  CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
  return SS_OK;
}
int ss_free_host(void *ptr) {
  if (ptr) CK(cudaFreeHost(ptr));
  return SS_OK;
}

int ss_pointer_device(const void *ptr, int32_t *device) {
  if (!device) return fail(SS_ERR_INVALID, "null argument");
  cudaPointerAttributes attr{};
  cudaError_t err = cudaPointerGetAttributes(&attr, ptr);
  if (err != cudaSuccess) { cudaGetLastError(); *device = -1; return fail(SS_ERR_CUDA, cudaGetErrorString(err)); }
  *device = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? attr.device : -1;
  return SS_OK;
}

int ss_create(const ss_config *cfg, const float *mapLx, const float *mapLy, const float *mapRx, const float *mapRy,
              const float *a1, const float *a2, const float *a3, ss_engine **out) {
  return create_lanes(cfg, nullptr, mapLx, mapLy, mapRx, mapRy, a1, a2, a3, out);
}
int ss_create_calibrated(const ss_config *cfg, const ss_calibration *cal, ss_engine **out) {
  if (!cal) return fail(SS_ERR_INVALID, "null calibration");
  return create_lanes(cfg, cal, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out);
}
int ss_destroy(ss_engine *e) {
  if (!e) return SS_OK;
  for (Core *c : e->lane) core_destroy(c);
  DeviceGuard g(e->device);
  if (e->pub) { cudaStreamSynchronize(e->pub); cudaStreamDestroy(e->pub); }
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  for (auto ev : e->ev_tail) if (ev) cudaEventDestroy(ev);
  delete e;
  return SS_OK;
}

#define SS_FRAME(call)                                                                                                 \
  if (!e) return fail(SS_ERR_INVALID, "null engine");                                                                  \
  DeviceGuard g(e->device);                                                                                            \
  const int ln = e->next_lane();                                                                                       \
  Core *c = e->lane[ln];                                                                                               \
  c->frame_id = e->frames;                                                                                             \
  int r = lane_begin(e, ln);                                                                                           \
  if (r) return r;                                                                                                     \
  r = (call);                                                                                                          \
  if (r) return r;                                                                                                     \
  return lane_end(e, ln);

int ss_compute_host_u8(ss_engine *e, const uint8_t *left, const uint8_t *right, const ss_bbox *bbox) {
  SS_FRAME(core_compute_host_u8(c, left, right, bbox))
}
int ss_compute_device_rgba_f32(ss_engine *e, const void *left, const void *right, const ss_bbox *bbox, void *stream) {
  SS_FRAME(core_compute_device_rgba_f32(c, left, right, bbox, user_stream(e, stream)))
}
int ss_compute_device_rgba_f32_pitched(ss_engine *e, const void *left, const void *right, size_t env_pitch_bytes,
                                       size_t row_pitch_bytes, const ss_bbox *bbox, void *stream) {
  SS_FRAME(core_compute_device_rgba_f32_pitched(c, left, right, env_pitch_bytes, row_pitch_bytes, bbox, user_stream(e, stream)))
}
int ss_compute_device_u8(ss_engine *e, const void *left, const void *right, const ss_bbox *bbox, void *stream) {
  SS_FRAME(core_compute_device_u8(c, left, right, bbox, user_stream(e, stream)))
}
int ss_submit_host_u8(ss_engine *e, const uint8_t *left, const uint8_t *right, const ss_bbox *bbox, float *out_host,
                      size_t capacity_bytes, uint64_t *ticket) {
  if (!ticket) return fail(SS_ERR_INVALID, "null argument");
  uint64_t sub = 0;
  const uint64_t mine = e ? e->frames + 1 : 0;
  SS_FRAME((r = core_submit_host_u8(c, left, right, bbox, out_host, capacity_bytes, &sub),
            r ? r : (e->tk[mine & 7] = {mine, sub, ln}, *ticket = mine, 0)))
}
int ss_wait_frame(ss_engine *e, uint64_t ticket) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (ticket == 0 || ticket > e->frames) return fail(SS_ERR_INVALID, "unknown frame ticket");
  const ss_engine::Ticket &t = e->tk[ticket & 7];
  if (t.pub != ticket) return SS_OK; // an old ticket: the lanes' in-flight limit has already waited for that frame
  return core_wait_frame(e->lane[t.lane], t.sub);
}
int ss_wait_stream(ss_engine *e, void *stream) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (stream == (void *)e->pub) return SS_OK;
  return core_wait_stream(e->lane[e->next_lane()], stream);
}
int ss_synchronize(ss_engine *e) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  for (int i = 0; i < e->nlanes; ++i) { int r = core_synchronize(e->lane[i]); if (r) return r; }
  DeviceGuard g(e->device);
  CK(cudaStreamSynchronize(e->pub));
  return SS_OK;
}

int ss_get_output_shape(const ss_engine *e, uint32_t *rows, uint32_t *cols) { return e ? core_get_output_shape(e->lane[0], rows, cols) : fail(SS_ERR_INVALID, "null argument"); }
int ss_get_input_shape(const ss_engine *e, uint32_t *rows, uint32_t *cols) { return e ? core_get_input_shape(e->lane[0], rows, cols) : fail(SS_ERR_INVALID, "null argument"); }
int ss_get_device(const ss_engine *e, int32_t *device) { return e ? core_get_device(e->lane[0], device) : fail(SS_ERR_INVALID, "null argument"); }
int ss_get_stream(const ss_engine *e, void **stream) {
  if (!e || !stream) return fail(SS_ERR_INVALID, "null argument");
  *stream = e->pub;
  return SS_OK;
}
int ss_get_lanes(const ss_engine *e, int32_t *lanes) {
  if (!e || !lanes) return fail(SS_ERR_INVALID, "null argument");
  *lanes = e->nlanes;
  return SS_OK;
}

#define SS_LAST(fn, ...) return e ? fn(e->lane[e->last], ##__VA_ARGS__) : fail(SS_ERR_INVALID, "null engine");
int ss_get_depth_host(ss_engine *e, float *out, size_t cap) { SS_LAST(core_get_depth_host, out, cap) }
int ss_get_depth_device(ss_engine *e, void **ptr) { SS_LAST(core_get_depth_device, ptr) }
int ss_get_point_cloud_host(ss_engine *e, float *out, size_t cap) { SS_LAST(core_get_point_cloud_host, out, cap) }
int ss_get_point_cloud_device(ss_engine *e, void **ptr) { SS_LAST(core_get_point_cloud_device, ptr) }
int ss_get_rgb_point_cloud_host(ss_engine *e, const void *rgba, float *out, size_t cap) { SS_LAST(core_get_rgb_point_cloud_host, rgba, out, cap) }
int ss_get_rgb_point_cloud_device(ss_engine *e, const void *rgba, void **ptr) { SS_LAST(core_get_rgb_point_cloud_device, rgba, ptr) }
int ss_get_stage_host(ss_engine *e, const char *name, int32_t index, void *out, size_t cap, size_t *bytes) {
  SS_LAST(core_get_stage_host, name, index, out, cap, bytes)
}
int ss_enqueue_point_cloud(ss_engine *e, const void *rgba_device, void **ptr) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  Core *c = e->lane[e->last];
  int r = core_enqueue_point_cloud(c, rgba_device, ptr);
  if (r) return r;
  DeviceGuard g(e->device); // the public stream completes behind the point-cloud kernel as well
  CK(cudaEventRecord(e->ev_join, c->stream));
  CK(cudaStreamWaitEvent(e->pub, e->ev_join, 0));
  return SS_OK;
}
int ss_get_launches_per_compute(ss_engine *e, int32_t *count) { SS_LAST(core_get_launches_per_compute, count) }

#define SS_ALL(fn, ...)                                                                                                \
  if (!e) return fail(SS_ERR_INVALID, "null engine");                                                                  \
  for (int i = 0; i < e->nlanes; ++i) { int r = fn(e->lane[i], ##__VA_ARGS__); if (r) return r; }                       \
  return SS_OK;
int ss_bind_output_host(ss_engine *e, float *out, size_t cap) { SS_ALL(core_bind_output_host, out, cap) }
int ss_set_ir_noise_parameters(ss_engine *e, float shape, float scale, float mu, float sigma) { SS_ALL(core_set_ir_noise_parameters, shape, scale, mu, sigma) }
int ss_set_penalties(ss_engine *e, int32_t p1, int32_t p2) { SS_ALL(core_set_penalties, p1, p2) }
int ss_set_census_window_size(ss_engine *e, int32_t w, int32_t h) { SS_ALL(core_set_census_window_size, w, h) }
int ss_set_matching_block_size(ss_engine *e, int32_t w, int32_t h) { SS_ALL(core_set_matching_block_size, w, h) }
int ss_set_uniqueness_ratio(ss_engine *e, int32_t u) { SS_ALL(core_set_uniqueness_ratio, u) }
int ss_set_lr_max_diff(ss_engine *e, int32_t d) { SS_ALL(core_set_lr_max_diff, d) }

int ss_set_profiling(ss_engine *e, int32_t enabled) {
  if (!e) return fail(SS_ERR_INVALID, "null engine");
  if (enabled && e->nlanes > 1) { int r = ss_synchronize(e); if (r) return r; } // lane 0 alone from here on
  e->profiling = enabled != 0;
  return core_set_profiling(e->lane[0], enabled);
}
int ss_get_stage_times(ss_engine *e, const char **names, float *ms, int32_t capacity, int32_t *count, int32_t *frames) {
  return e ? core_get_stage_times(e->lane[0], names, ms, capacity, count, frames) : fail(SS_ERR_INVALID, "null engine");
}

} // extern "C"
