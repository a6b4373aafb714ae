// pybind.cpp -- Python module `sapien_b200._simsense_b200`, the pybind layer over the C ABI
// (include/ss_b200.h).  Mirrors the reference's binding class DepthSensorEnginePython
// (python/pybind/simsense.cpp:52-179) and its CudaArray hand-off type
// (python/pybind/sapien.cpp:271-349, src/array.cpp:121-149): same constructor argument order,
// method names, keyword names and error types, so replacing
//   from ..pysapien.simsense import DepthSensorEngine        (simsense_component.py:18)
// by  from sapien_b200.simsense import DepthSensorEngine  is the whole integration.
// No torch types and no CUDA runtime calls here: everything goes through libss_b200.so.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ss_b200.h"

namespace py = pybind11;
using namespace pybind11::literals;

// ---- minimal DLPack ABI (dlpack.h v0.8 layout; the reference vendors 3rd_party/dlpack) --------
extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void *data; DLDevice device; int32_t ndim; DLDataType dtype;
  int64_t *shape; int64_t *strides; uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor; void *manager_ctx; void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;
}
static constexpr int kDLCUDA = 2;

namespace {

[[noreturn]] void raise_status(int status) {
  const std::string msg = ss_last_error();
  if (status == SS_ERR_INVALID) throw py::type_error(msg);
  throw std::runtime_error(msg);
}
inline void check(int status) { if (status != SS_OK) raise_status(status); }

int item_size(const std::string &t) { // "f4", "<f4", "|u1" ...
  size_t i = 0;
  while (i < t.size() && !isdigit((unsigned char)t[i])) ++i;
  return i < t.size() ? std::stoi(t.substr(i)) : 1;
}
char type_kind(const std::string &t) {
  for (char ch : t) if (isalpha((unsigned char)ch)) return ch;
  return '?';
}

// Non-owning (optionally owner-pinning) view of device memory: sapien::CudaArrayHandle
// (include/sapien/array.h:29-44) with the same Python attribute set.
struct CudaArray {
  std::vector<int64_t> shape, strides; // strides in BYTES, like the reference
  std::string type;
  int cuda_id = -1;
  void *ptr = nullptr;
  uintptr_t stream = 0; // producer stream of a __cuda_array_interface__ v3 object (0: none given)
  py::object owner; // keeps the producer alive (the reference does not; harmless extension)

  bool contiguous() const {
    int64_t expect = item_size(type);
    for (int i = (int)shape.size() - 1; i >= 0; --i) {
      if (shape[i] != 1 && strides[i] != expect) return false;
      expect *= shape[i];
    }
    return true;
  }
};

CudaArray from_object(py::object obj) {
  if (py::isinstance<CudaArray>(obj)) return obj.cast<CudaArray>();
  if (!py::hasattr(obj, "__cuda_array_interface__"))
    throw py::type_error("expected an object with __cuda_array_interface__");
  auto iface = obj.attr("__cuda_array_interface__").cast<py::dict>();
  CudaArray a;
  a.shape = iface["shape"].cast<py::tuple>().cast<std::vector<int64_t>>();
  a.type = iface["typestr"].cast<std::string>();
  if (iface.contains("strides") && !iface["strides"].is_none()) {
    a.strides = iface["strides"].cast<py::tuple>().cast<std::vector<int64_t>>();
  } else {
    a.strides.resize(a.shape.size());
    int64_t s = item_size(a.type);
    for (int i = (int)a.shape.size() - 1; i >= 0; --i) { a.strides[i] = s; s *= a.shape[i]; }
  }
  auto data = iface["data"].cast<py::tuple>();
  a.ptr = reinterpret_cast<void *>(data[0].cast<uintptr_t>());
  int32_t dev = -1;
  if (a.ptr && ss_pointer_device(a.ptr, &dev) != SS_OK) dev = -1;
  a.cuda_id = dev;
  if (iface.contains("stream") && !iface["stream"].is_none()) // v3: 1 = legacy default, 2 = per-thread default, else a handle
    a.stream = iface["stream"].cast<uintptr_t>();
  a.owner = obj;
  return a;
}

void dl_deleter(DLManagedTensor *self) {
  delete[] self->dl_tensor.shape;
  delete[] self->dl_tensor.strides;
  if (self->manager_ctx) {
    py::gil_scoped_acquire gil;
    delete static_cast<py::object *>(self->manager_ctx);
  }
  delete self;
}
void capsule_destructor(PyObject *cap) {
  // still named "dltensor" => nobody consumed it => we own the tensor (DLPack protocol)
  if (PyCapsule_IsValid(cap, "dltensor")) {
    auto *t = static_cast<DLManagedTensor *>(PyCapsule_GetPointer(cap, "dltensor"));
    if (t && t->deleter) t->deleter(t);
  } else {
    PyErr_Clear();
  }
}
py::object to_dlpack(const CudaArray &a, py::object self) {
  auto *t = new DLManagedTensor();
  t->manager_ctx = new py::object(self); // the capsule consumer keeps the CudaArray alive
  t->deleter = &dl_deleter;
  t->dl_tensor.data = a.ptr;
  t->dl_tensor.device = {kDLCUDA, a.cuda_id};
  t->dl_tensor.ndim = (int32_t)a.shape.size();
  const char k = type_kind(a.type);
  const int bytes = item_size(a.type);
  t->dl_tensor.dtype = {(uint8_t)(k == 'f' ? 2 : (k == 'u' ? 1 : 0)), (uint8_t)(8 * bytes), 1};
  t->dl_tensor.shape = new int64_t[a.shape.size()];
  t->dl_tensor.strides = new int64_t[a.shape.size()];
  for (size_t i = 0; i < a.shape.size(); ++i) {
    t->dl_tensor.shape[i] = a.shape[i];
    t->dl_tensor.strides[i] = a.strides[i] / bytes; // element strides (src/array.cpp:144)
  }
  t->dl_tensor.byte_offset = 0;
  return py::reinterpret_steal<py::object>(PyCapsule_New(t, "dltensor", &capsule_destructor));
}

const float *f32_plane(const py::array_t<float, py::array::c_style | py::array::forcecast> &a,
                       size_t expect, const char *name, bool required) {
  if (a.size() == 0 && !required) return nullptr;
  if ((size_t)a.size() != expect)
    throw std::runtime_error(std::string(name) + " must have rows*cols elements");
  return a.data();
}

class DepthSensorEngine {
public:
  using F32 = py::array_t<float, py::array::c_style | py::array::forcecast>;
  using U8 = py::array_t<uint8_t, py::array::c_style>;

  DepthSensorEngine(uint32_t rows, uint32_t cols, uint32_t rgbRows, uint32_t rgbCols, float focalLen,
                    float baselineLen, float minDepth, float maxDepth, uint64_t noiseSeed,
                    float speckleShape, float speckleScale, float gaussianMu, float gaussianSigma,
                    bool rectified, uint8_t censusWidth, uint8_t censusHeight, uint32_t maxDisp, uint8_t bfWidth,
                    uint8_t bfHeight, uint8_t p1, uint8_t p2, uint8_t uniqRatio, int lrMaxDiff, uint8_t mfSize, F32 mapLx,
                    F32 mapLy, F32 mapRx, F32 mapRy, F32 a1, F32 a2, F32 a3, float b1, float b2,
                    float b3, bool dilation, float mainFx, float mainFy, float mainSkew, float mainCx,
                    float mainCy, int device, int batch, bool keepStages, bool batched, py::object calibration, int lanes) {
    ss_config c{};
    c.rows = rows; c.cols = cols; c.rgb_rows = rgbRows; c.rgb_cols = rgbCols;
    c.focal_len = focalLen; c.baseline_len = baselineLen; c.min_depth = minDepth; c.max_depth = maxDepth;
    c.ir_noise_seed = noiseSeed; c.speckle_shape = speckleShape; c.speckle_scale = speckleScale;
    c.gaussian_mu = gaussianMu; c.gaussian_sigma = gaussianSigma; c.rectified = rectified;
    c.census_width = censusWidth; c.census_height = censusHeight; c.max_disp = (int32_t)std::min<uint32_t>(maxDisp, 1u << 20);
    c.bf_width = bfWidth; c.bf_height = bfHeight; c.p1 = p1; c.p2 = p2; c.uniq_ratio = uniqRatio;
    c.lr_max_diff = lrMaxDiff; c.mf_size = mfSize; c.b1 = b1; c.b2 = b2; c.b3 = b3;
    c.dilation = dilation; c.main_fx = mainFx; c.main_fy = mainFy; c.main_skew = mainSkew;
    c.main_cx = mainCx; c.main_cy = mainCy; c.registration = 1;
    c.device = device; c.batch = batch; c.keep_stages = keepStages; c.lanes = lanes;
    const size_t n = (size_t)rows * cols;
    if (!calibration.is_none()) {
      // extension: calibration=(reg_m 3x3, rect_inv_left 3x3, rect_inv_right 3x3, (fx, fy, cx, cy)) float64 -- the planes
      // are evaluated in the kernels; the seven plane arguments are ignored (pass empty arrays)
      using F64 = py::array_t<double, py::array::c_style | py::array::forcecast>;
      auto t = calibration.cast<py::tuple>();
      if (t.size() != 4) throw py::type_error("calibration must be (reg_m, rect_inv_left, rect_inv_right, (fx, fy, cx, cy))");
      ss_calibration cal{};
      auto m3 = [](py::handle h, double (&dst)[9], const char *name) {
        auto a = h.cast<F64>();
        if (a.size() != 9) throw py::type_error(std::string(name) + " must be a 3x3 matrix");
        std::memcpy(dst, a.data(), sizeof(dst));
      };
      m3(t[0], cal.reg_m, "reg_m");
      m3(t[1], cal.rect_inv_left, "rect_inv_left");
      m3(t[2], cal.rect_inv_right, "rect_inv_right");
      auto k = t[3].cast<F64>();
      if (k.size() != 4) throw py::type_error("ir camera must be (fx, fy, cx, cy)");
      cal.ir_fx = k.data()[0]; cal.ir_fy = k.data()[1]; cal.ir_cx = k.data()[2]; cal.ir_cy = k.data()[3];
      check(ss_create_calibrated(&c, &cal, &e_));
    } else {
      check(ss_create(&c, f32_plane(mapLx, n, "map_lx", !rectified), f32_plane(mapLy, n, "map_ly", !rectified),
                      f32_plane(mapRx, n, "map_rx", !rectified), f32_plane(mapRy, n, "map_ry", !rectified),
                      f32_plane(a1, n, "a1", true), f32_plane(a2, n, "a2", true), f32_plane(a3, n, "a3", true),
                      &e_));
    }
    rows_ = rows; cols_ = cols; batch_ = batch; lead_ = (batch > 1 || batched) ? 1 : 0;
    check(ss_get_output_shape(e_, &orows_, &ocols_));
    check(ss_get_device(e_, &device_));
  }
  ~DepthSensorEngine() { if (e_) ss_destroy(e_); }
  DepthSensorEngine(const DepthSensorEngine &) = delete;

  // python/pybind/simsense.cpp:75-81
  void computeHost(U8 left, U8 right, bool bbox, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    checkHostShape(left, right);
    ss_bbox bb{bbox ? 1 : 0, x, y, w, h};
    const uint8_t *l = left.data(), *r = right.data();
    // A strict compute(l, r) + get_ndarray() caller: the frame is delivered straight into a page-locked array from
    // a small pool and get_ndarray() returns THAT array -- no staging copy, no fresh allocation.  An array is reused
    // only when nobody but the pool references it any more, so results a caller still holds are never overwritten.
    pool_last_ = -1;
    pool_given_ = false;
    float *dst = nullptr;
    const size_t bytes = (size_t)std::max(batch_, 1) * orows_ * ocols_ * sizeof(float);
    if (bound_.is_none() && bytes <= (64u << 20)) {
      int slot = -1;
      for (size_t i = 0; i < pool_.size(); ++i)
        if (pool_[i].ref_count() == 1) { slot = (int)i; break; }
      if (slot < 0 && pool_.size() < 4) {
        void *p = nullptr;
        if (ss_alloc_host(bytes, &p) == SS_OK) {
          py::capsule owner(p, [](void *q) { ss_free_host(q); });
          std::vector<py::ssize_t> shape{(py::ssize_t)orows_, (py::ssize_t)ocols_};
          if (lead_) shape.insert(shape.begin(), batch_);
          pool_.push_back(py::array_t<float>(shape, static_cast<float *>(p), owner));
          slot = (int)pool_.size() - 1;
        }
      }
      if (slot >= 0) { dst = pool_[slot].mutable_data(); pool_last_ = slot; }
    }
    py::gil_scoped_release nogil;
    int st;
    if (dst) {
      uint64_t ticket = 0;
      st = ss_submit_host_u8(e_, l, r, &bb, dst, bytes, &ticket);
      if (!st) st = ss_wait_frame(e_, ticket);
    } else {
      st = ss_compute_host_u8(e_, l, r, &bb);
    }
    if (st) { py::gil_scoped_acquire gil; pool_last_ = -1; raise_status(st); }
  }
  // extension: asynchronous host frames (ss_submit_host_u8 / ss_wait_frame).  The arrays are kept alive until waited for.
  uint64_t submitHost(U8 left, U8 right, py::object out_arg, bool bbox, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    checkHostShape(left, right);
    float *dst = nullptr;
    size_t cap = 0;
    if (!out_arg.is_none()) {
      auto out = out_arg.cast<py::array_t<float, py::array::c_style>>();
      if (out.ptr() != out_arg.ptr()) throw py::type_error("out must be a C-contiguous float32 ndarray");
      dst = out.mutable_data();
      cap = (size_t)out.nbytes();
    }
    ss_bbox bb{bbox ? 1 : 0, x, y, w, h};
    uint64_t ticket = 0;
    pool_last_ = -1;
    check(ss_submit_host_u8(e_, left.data(), right.data(), &bb, dst, cap, &ticket));
    inflight_[ticket] = py::make_tuple(left, right, out_arg);
    while (inflight_.size() > 2) inflight_.erase(inflight_.begin()); // older frames were waited for inside submit
    return ticket;
  }
  void waitFrame(uint64_t ticket) {
    int st;
    { py::gil_scoped_release nogil; st = ss_wait_frame(e_, ticket); }
    check(st);
    inflight_.erase(ticket);
  }
  // python/pybind/simsense.cpp:83-99
  void computeCuda(py::object leftObj, py::object rightObj, bool bbox, uint32_t x, uint32_t y,
                   uint32_t w, uint32_t h, py::object stream, bool sync) {
    CudaArray left = from_object(leftObj), right = from_object(rightObj);
    pool_last_ = -1;
    const size_t lead = lead_;
    if (left.shape.size() < 2 + lead || right.shape.size() < 2 + lead)
      throw std::runtime_error("Input image size different from initiated");
    if (left.shape[lead] != right.shape[lead] || left.shape[lead + 1] != right.shape[lead + 1])
      throw std::runtime_error("Both images must have the same size");
    if (left.shape[lead] != rows_ || left.shape[lead + 1] != cols_ || (lead && left.shape[0] != batch_))
      throw std::runtime_error("Input image size different from initiated");
    const char k = type_kind(left.type);
    const bool is_f4 = k == 'f' && item_size(left.type) == 4 && type_kind(right.type) == 'f' && item_size(right.type) == 4;
    const bool is_u1 = k == 'u' && item_size(left.type) == 1 && type_kind(right.type) == 'u' && item_size(right.type) == 1;
    if (!is_f4 && !is_u1) throw std::runtime_error("Input data type must be float");
    if (is_f4 && (left.shape.size() != 3 + lead || left.shape.back() != 4 || right.shape.size() != 3 + lead || right.shape.back() != 4))
      throw std::runtime_error("float input must be an RGBA image [H, W, 4]");
    // Packed arrays, or pitched float RGBA views (a window of a BatchedCamera buffer, a sliced tensor): pixels
    // must be 4 consecutive floats; row and environment pitches are free but common to both images.
    bool pitched = false;
    size_t env_pitch = 0, row_pitch = 0;
    if (!left.contiguous() || !right.contiguous()) {
      const size_t nd = left.shape.size();
      bool ok = is_f4 && left.strides == right.strides && left.strides[nd - 1] == 4 && left.strides[nd - 2] == 16 &&
                left.strides[nd - 3] >= (int64_t)cols_ * 16 && (!lead || left.strides[0] > 0);
      if (!ok) throw std::runtime_error("input CUDA arrays must be C-contiguous, or float RGBA views with 16-byte pixels and equal pitches");
      pitched = true;
      row_pitch = (size_t)left.strides[nd - 3];
      env_pitch = lead ? (size_t)left.strides[0] : row_pitch * rows_;
    }
    if ((left.cuda_id >= 0 && left.cuda_id != device_) || (right.cuda_id >= 0 && right.cuda_id != device_))
      throw std::runtime_error("input CUDA arrays live on a different device than the engine");
    ss_bbox bb{bbox ? 1 : 0, x, y, w, h};
    // stream=None does NOT mean "inputs are ready": the frame is ordered after the legacy default stream
    // (where torch / cupy enqueue unless told otherwise) -- the reference orders by a device-wide sync at
    // the start of the frame (core.cu:547).  Producer streams announced by the inputs themselves
    // (__cuda_array_interface__ v3 `stream`) are honoured as well.  Pass stream=engine.cuda_stream to
    // state that the inputs are complete and no ordering is wanted.
    static constexpr uintptr_t kLegacy = 1;
    void *s = reinterpret_cast<void *>(stream.is_none() ? kLegacy : stream.cast<uintptr_t>());
    for (uintptr_t ps : {left.stream, right.stream})
      if (ps && reinterpret_cast<void *>(ps) != s) check(ss_wait_stream(e_, reinterpret_cast<void *>(ps)));
    py::gil_scoped_release nogil;
    int st = pitched ? ss_compute_device_rgba_f32_pitched(e_, left.ptr, right.ptr, env_pitch, row_pitch, &bb, s)
             : is_f4 ? ss_compute_device_rgba_f32(e_, left.ptr, right.ptr, &bb, s)
                     : ss_compute_device_u8(e_, left.ptr, right.ptr, &bb, s);
    if (!st && sync) st = ss_synchronize(e_);
    if (st) { py::gil_scoped_acquire gil; raise_status(st); }
  }

  py::array_t<float> getNdarray(py::object out_arg) {
    if (!out_arg.is_none()) { // extension: write into a caller-provided (e.g. pinned) float32 array
      auto out = out_arg.cast<py::array_t<float, py::array::c_style>>();
      if (out.ptr() != out_arg.ptr()) throw py::type_error("out must be a C-contiguous float32 ndarray");
      float *dst = out.mutable_data();
      const size_t cap = (size_t)out.nbytes();
      int st;
      { py::gil_scoped_release nogil; st = ss_get_depth_host(e_, dst, cap); }
      check(st);
      return out;
    }
    if (pool_last_ >= 0) { // delivered into a pool array by the last compute(host): hand it out (a copy on repeated calls)
      if (!pool_given_) { pool_given_ = true; return pool_[pool_last_]; }
      return py::array_t<float>(py::array(pool_[pool_last_]).attr("copy")());
    }
    auto out = make_out({(py::ssize_t)orows_, (py::ssize_t)ocols_});
    float *dst = out.mutable_data();
    const size_t cap = (size_t)out.nbytes();
    int st;
    { py::gil_scoped_release nogil; st = ss_get_depth_host(e_, dst, cap); }
    check(st);
    return out;
  }
  // extension: every following compute() streams its depth map into `out` (pinned float32 ndarray of
  // the get_ndarray shape) while the last aggregation pass is still running; get_ndarray(out=out)
  // then only waits.  None unbinds.  The engine keeps a reference to the array.
  void bindOutput(py::object out_arg) {
    if (out_arg.is_none()) { check(ss_bind_output_host(e_, nullptr, 0)); bound_ = py::none(); return; }
    auto out = out_arg.cast<py::array_t<float, py::array::c_style>>();
    if (out.ptr() != out_arg.ptr()) throw py::type_error("out must be a C-contiguous float32 ndarray");
    check(ss_bind_output_host(e_, out.mutable_data(), (size_t)out.nbytes()));
    bound_ = out_arg;
  }
  CudaArray getCuda() {
    void *p = nullptr;
    check(ss_get_depth_device(e_, &p));
    return view(p, {(int64_t)orows_, (int64_t)ocols_});
  }
  py::array_t<float> getPointCloudNdarray() {
    auto out = make_out({(py::ssize_t)orows_ * ocols_, 3});
    check(ss_get_point_cloud_host(e_, out.mutable_data(), (size_t)out.nbytes()));
    return out;
  }
  CudaArray getPointCloudCuda(bool sync) {
    void *p = nullptr;
    check(sync ? ss_get_point_cloud_device(e_, &p) : ss_enqueue_point_cloud(e_, nullptr, &p));
    return view(p, {(int64_t)orows_ * ocols_, 3});
  }
  py::array_t<float> getRgbPointCloudNdarray(py::object rgba) {
    CudaArray a = checkedRgba(rgba);
    auto out = make_out({(py::ssize_t)orows_ * ocols_, 6});
    check(ss_get_rgb_point_cloud_host(e_, a.ptr, out.mutable_data(), (size_t)out.nbytes()));
    return out;
  }
  CudaArray getRgbPointCloudCuda(py::object rgba, bool sync) {
    CudaArray a = checkedRgba(rgba);
    void *p = nullptr;
    check(sync ? ss_get_rgb_point_cloud_device(e_, a.ptr, &p) : ss_enqueue_point_cloud(e_, a.ptr, &p));
    return view(p, {(int64_t)orows_ * ocols_, 6});
  }

  void setIrNoise(float a, float b, float c, float d) { check(ss_set_ir_noise_parameters(e_, a, b, c, d)); }
  void setCensus(int w, int h) { check(ss_set_census_window_size(e_, w, h)); }
  void setBlock(int w, int h) { check(ss_set_matching_block_size(e_, w, h)); }
  void setPenalties(int p1, int p2) { check(ss_set_penalties(e_, p1, p2)); }
  void setUniq(int u) { check(ss_set_uniqueness_ratio(e_, u)); }
  void setLr(int d) { check(ss_set_lr_max_diff(e_, d)); }
  void synchronize() { int st; { py::gil_scoped_release nogil; st = ss_synchronize(e_); } check(st); }
  void waitStream(uintptr_t stream) { check(ss_wait_stream(e_, reinterpret_cast<void *>(stream))); }
  void setProfiling(bool on) { check(ss_set_profiling(e_, on)); }
  py::dict stageTimes() {
    const char *names[64]; float ms[64]; int32_t n = 0, frames = 0;
    check(ss_get_stage_times(e_, names, ms, 64, &n, &frames));
    py::dict d;
    for (int i = 0; i < n && i < 64; ++i) d[py::str(names[i])] = ms[i] / (frames > 0 ? frames : 1);
    d["frames"] = frames;
    return d;
  }
  int launches() { int32_t n = 0; check(ss_get_launches_per_compute(e_, &n)); return n; }
  py::object getStage(const std::string &name, int index) {
    size_t bytes = 0;
    std::vector<char> probe(1);
    int st = ss_get_stage_host(e_, name.c_str(), index, probe.data(), 0, &bytes);
    if (st != SS_OK && bytes == 0) raise_status(st);
    std::vector<char> buf(bytes);
    check(ss_get_stage_host(e_, name.c_str(), index, buf.data(), bytes, &bytes));
    const char *dt = (name == "im0" || name == "im1") ? "uint8"
                   : (name == "census0" || name == "census1") ? "uint32"
                   : ((name.rfind("disp_", 0) == 0 && name != "disp_right") || name == "depth") ? "float32"
                   : "uint16";
    py::array_t<uint8_t> raw((py::ssize_t)bytes);
    std::memcpy(raw.mutable_data(), buf.data(), bytes);
    py::object arr = raw.attr("view")(py::str(dt));
    return arr;
  }
  uint32_t inRows() const { return rows_; }
  uint32_t inCols() const { return cols_; }
  uint32_t outRows() const { return orows_; }
  uint32_t outCols() const { return ocols_; }
  int device() const { return device_; }
  int batch() const { return batch_; }
  int lanes() const { int32_t n = 1; check(ss_get_lanes(e_, &n)); return n; }
  uintptr_t cudaStream() const { void *st = nullptr; check(ss_get_stream(e_, &st)); return reinterpret_cast<uintptr_t>(st); }

private:
  void checkHostShape(const U8 &l, const U8 &r) {
    const size_t lead = lead_;
    if (l.ndim() != (py::ssize_t)(2 + lead) || r.ndim() != (py::ssize_t)(2 + lead))
      throw std::runtime_error("Input image size different from initiated");
    for (int i = 0; i < l.ndim(); ++i)
      if (l.shape(i) != r.shape(i)) throw std::runtime_error("Both images must have the same size");
    if ((uint32_t)l.shape(lead) != rows_ || (uint32_t)l.shape(lead + 1) != cols_ || (lead && l.shape(0) != batch_))
      throw std::runtime_error("Input image size different from initiated");
  }
  CudaArray checkedRgba(py::object rgba) {
    CudaArray a = from_object(rgba);
    const size_t lead = lead_;
    if (type_kind(a.type) != 'f' || item_size(a.type) != 4 || a.shape.size() != 3 + lead ||
        a.shape[lead] != orows_ || a.shape[lead + 1] != ocols_ || a.shape[lead + 2] != 4 || !a.contiguous())
      throw std::runtime_error("rgba must be a contiguous float32 CUDA array of shape [out_rows, out_cols, 4]");
    return a;
  }
  py::array_t<float> make_out(std::vector<py::ssize_t> shape) {
    if (lead_) shape.insert(shape.begin(), batch_);
    return py::array_t<float>(shape);
  }
  CudaArray view(void *p, std::vector<int64_t> shape) {
    if (lead_) shape.insert(shape.begin(), batch_);
    CudaArray a;
    a.shape = shape;
    a.strides.resize(shape.size());
    int64_t s = 4;
    for (int i = (int)shape.size() - 1; i >= 0; --i) { a.strides[i] = s; s *= shape[i]; }
    a.type = "f4"; // python/pybind/simsense.cpp:106
    a.cuda_id = device_;
    a.ptr = p;
    return a;
  }
  ss_engine *e_ = nullptr;
  uint32_t rows_ = 0, cols_ = 0, orows_ = 0, ocols_ = 0;
  int32_t device_ = 0;
  int batch_ = 1;
  size_t lead_ = 0; // 1: inputs and outputs carry a leading environment dimension (batch > 1, or batched=True)
  py::object bound_ = py::none();
  std::map<uint64_t, py::object> inflight_;
  std::vector<py::array_t<float>> pool_; // page-locked result arrays of the strict host path
  int pool_last_ = -1;                   // pool array that holds the last host frame (-1: none)
  bool pool_given_ = false;
};

} // namespace

PYBIND11_MODULE(_simsense_b200, m) {
  m.doc() = "B200-native drop-in for sapien.pysapien.simsense (DepthSensorEngine) + CudaArray";
  m.def("version", []() { return std::string(ss_version()); });

  py::class_<CudaArray>(m, "CudaArray")
      .def(py::init([](py::object obj) { return from_object(obj); }), py::arg("data"))
      .def_property_readonly("shape", [](CudaArray &a) { return a.shape; })
      .def_property_readonly("strides", [](CudaArray &a) { return a.strides; })
      .def_readonly("cuda_id", &CudaArray::cuda_id)
      .def_readonly("typestr", &CudaArray::type)
      .def_property_readonly("ptr", [](CudaArray &a) { return reinterpret_cast<intptr_t>(a.ptr); })
      .def_property_readonly("__cuda_array_interface__",
                             [](CudaArray &a) {
                               return py::dict("shape"_a = py::tuple(py::cast(a.shape)),
                                               "strides"_a = py::tuple(py::cast(a.strides)),
                                               "typestr"_a = a.type,
                                               "data"_a = py::make_tuple(reinterpret_cast<intptr_t>(a.ptr), false),
                                               "version"_a = 2);
                             })
      .def("torch",
           [](py::object self) {
             CudaArray &a = self.cast<CudaArray &>();
             CudaArray b = a; // torch has no unsigned types beyond uint8 (sapien.cpp:321-325)
             if (b.type != "u1" && type_kind(b.type) == 'u') {
               b.type = std::string("i") + std::to_string(item_size(b.type));
             }
             b.owner = self;
             auto as_tensor = py::module_::import("torch").attr("as_tensor");
             const std::string dev = a.cuda_id >= 0 ? "cuda:" + std::to_string(a.cuda_id) : "cuda";
             return as_tensor(py::cast(b), "device"_a = dev);
           })
      .def("dlpack", [](py::object self) { return to_dlpack(self.cast<CudaArray &>(), self); })
      .def("__dlpack__", [](py::object self, py::kwargs) { return to_dlpack(self.cast<CudaArray &>(), self); })
      .def("__dlpack_device__", [](CudaArray &a) { return py::make_tuple(kDLCUDA, a.cuda_id); })
      .def("jax", [](py::object self) {
        auto from_dlpack = py::module_::import("jax").attr("dlpack").attr("from_dlpack");
        return from_dlpack(to_dlpack(self.cast<CudaArray &>(), self));
      });

  using E = DepthSensorEngine;
  py::class_<E>(m, "DepthSensorEngine")
      // the ten small integers are uint8_t like python/pybind/simsense.cpp:54-73 (lr_max_diff stays int: the
      // Python layer documents -1 as "off", which the reference's uint8_t wraps to 255)
      .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t, float, float, float, float, uint64_t, float,
                    float, float, float, bool, uint8_t, uint8_t, uint32_t, uint8_t, uint8_t, uint8_t, uint8_t, uint8_t, int, uint8_t, E::F32,
                    E::F32, E::F32, E::F32, E::F32, E::F32, E::F32, float, float, float, bool, float,
                    float, float, float, float, int, int, bool, bool, py::object, int>(),
           "rows"_a, "cols"_a, "rgb_rows"_a, "rgb_cols"_a, "focal_len"_a, "baseline_len"_a,
           "min_depth"_a, "max_depth"_a, "ir_noise_seed"_a, "speckle_shape"_a, "speckle_scale"_a,
           "gaussian_mu"_a, "gaussian_sigma"_a, "rectified"_a, "census_width"_a, "census_height"_a,
           "max_disp"_a, "bf_width"_a, "bf_height"_a, "p1"_a, "p2"_a, "uniq_ratio"_a, "lr_max_diff"_a,
           "mf_size"_a, "map_lx"_a, "map_ly"_a, "map_rx"_a, "map_ry"_a, "a1"_a, "a2"_a, "a3"_a, "b1"_a,
           "b2"_a, "b3"_a, "dilation"_a, "main_fx"_a, "main_fy"_a, "main_skew"_a, "main_cx"_a,
           "main_cy"_a, "device"_a = -1, "batch"_a = 1, "keep_stages"_a = false, "batched"_a = false, "calibration"_a = py::none(), "lanes"_a = 0)
      .def("compute", &E::computeHost, "left_array"_a, "right_array"_a, "bbox"_a = false,
           "bbox_start_x"_a = 0, "bbox_start_y"_a = 0, "bbox_width"_a = 0, "bbox_height"_a = 0)
      .def("compute", &E::computeCuda, "left_cuda"_a, "right_cuda"_a, "bbox"_a = false,
           "bbox_start_x"_a = 0, "bbox_start_y"_a = 0, "bbox_width"_a = 0, "bbox_height"_a = 0,
           "stream"_a = py::none(), "sync"_a = true)
      .def("get_ndarray", &E::getNdarray, "out"_a = py::none())
      .def("bind_output", &E::bindOutput, "out"_a)
      .def("get_cuda", &E::getCuda)
      .def("get_point_cloud_cuda", &E::getPointCloudCuda, "sync"_a = true)
      .def("get_point_cloud_ndarray", &E::getPointCloudNdarray)
      .def("get_rgb_point_cloud_ndarray", &E::getRgbPointCloudNdarray)
      .def("get_rgb_point_cloud_cuda", &E::getRgbPointCloudCuda, "rgba_cuda"_a, "sync"_a = true)
      .def("set_ir_noise_parameters", &E::setIrNoise)
      .def("set_census_window_size", &E::setCensus)
      .def("set_matching_block_size", &E::setBlock)
      .def("set_penalties", &E::setPenalties)
      .def("set_uniqueness_ratio", &E::setUniq)
      .def("set_lr_max_diff", &E::setLr)
      // ---- extensions ----
      .def("submit", &E::submitHost, "left_array"_a, "right_array"_a, "out"_a = py::none(), "bbox"_a = false,
           "bbox_start_x"_a = 0, "bbox_start_y"_a = 0, "bbox_width"_a = 0, "bbox_height"_a = 0)
      .def("wait", &E::waitFrame, "ticket"_a)
      .def("synchronize", &E::synchronize)
      .def("wait_stream", &E::waitStream, "stream"_a)
      .def("set_profiling", &E::setProfiling)
      .def("get_stage_times", &E::stageTimes)
      .def("get_launches_per_compute", &E::launches)
      .def("get_stage", &E::getStage, "name"_a, "index"_a = 0)
      .def_property_readonly("input_rows", &E::inRows)
      .def_property_readonly("input_cols", &E::inCols)
      .def_property_readonly("output_rows", &E::outRows)
      .def_property_readonly("output_cols", &E::outCols)
      .def_property_readonly("cuda_id", &E::device)
      .def_property_readonly("batch", &E::batch)
      .def_property_readonly("lanes", &E::lanes)
      .def_property_readonly("cuda_stream", &E::cudaStream);
}
