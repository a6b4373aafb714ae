/*
 * ss_b200.h -- C ABI of the B200-native stereo depth engine (libss_b200.so).
 *
 * This is the drop-in boundary for SAPIEN's GPU stereo depth path.  Each entry point names the
 * reference interface it replaces (paths relative to the SAPIEN tree):
 *   E = 3rd_party/simsense/include/simsense/core.h   (class simsense::DepthSensorEngine)
 *   C = 3rd_party/simsense/src/core.cu
 *   P = python/pybind/simsense.cpp                   (class DepthSensorEnginePython)
 *
 * Conventions: plain pointers and sizes only; every function returns an ss_status (0 = OK) and
 * never terminates the process (the reference calls exit() on CUDA errors, C:30-39);
 * ss_last_error() gives the message for the calling thread.  An engine is bound to one CUDA
 * device at creation and owns all of its buffers; device getters return BORROWED pointers that
 * stay valid until the next compute on that engine (same lifetime rule as P:101-111).  All work is
 * enqueued on the engine's stream; host-returning calls synchronise that stream only (the
 * reference issues 13-16 device-wide cudaDeviceSynchronize per frame, C:547-786).
 *
 * Batching is an extension: an engine created with batch = N processes N independent stereo
 * pairs ("environments") per compute call; every image argument then has a leading dimension N
 * and every output too.  batch = 1 is exactly the reference API.
 */
#ifndef SS_B200_H
#define SS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss_engine ss_engine;

typedef enum {
  SS_OK = 0,
  SS_ERR_INVALID = 1,      /* bad argument / config (Python layer raises TypeError / RuntimeError) */
  SS_ERR_NOT_COMPUTED = 2, /* "No computed data stored" (C:348-350)                               */
  SS_ERR_CUDA = 3,         /* CUDA runtime error, message in ss_last_error()                      */
  SS_ERR_NO_DEVICE = 4     /* no usable sm_100 device: there is NO CPU fallback                   */
} ss_status;

/* Scalar configuration = the scalar arguments of the registration constructor E:43-55 / P:54-73,
 * in the same order, plus device/batch/flags. */
typedef struct {
  uint32_t rows, cols;          /* IR image size                                              */
  uint32_t rgb_rows, rgb_cols;  /* RGB (output) size when registration != 0                   */
  float focal_len;              /* px                                                          */
  float baseline_len;           /* m                                                           */
  float min_depth, max_depth;   /* m                                                           */
  uint64_t ir_noise_seed;
  float speckle_shape, speckle_scale, gaussian_mu, gaussian_sigma; /* shape<=0: noise off    */
  int32_t rectified;            /* 0: apply the rectification maps                            */
  int32_t census_width, census_height;
  int32_t max_disp;
  int32_t bf_width, bf_height;  /* matching block                                             */
  int32_t p1, p2;
  int32_t uniq_ratio;
  int32_t lr_max_diff;          /* 255 disables the LR check (C:709)                          */
  int32_t mf_size;              /* 1,3,5,7                                                     */
  float b1, b2, b3;             /* registration translation K_rgb*t                           */
  int32_t dilation;
  float main_fx, main_fy, main_skew, main_cx, main_cy; /* point-cloud intrinsics              */
  int32_t registration;         /* 1: E:43-55 (the only ctor Python can reach); 0: E:33-41    */
  /* ---- extensions ---- */
  int32_t device;               /* CUDA device ordinal; -1 = current device                   */
  int32_t batch;                /* number of stereo pairs per compute call (>=1)              */
  int32_t keep_stages;          /* 1: also materialise L3/LAll so ss_get_stage can read them  */
  int32_t lanes;                /* 0: automatic (3 for batch == 1 without keep_stages, else 1); 1, 2 or 3: see ss_get_lanes */
} ss_config;

/* E:43-55 + C:178-306.  map* / a* are host float32 [rows*cols] arrays (maps may be NULL when
 * rectified, a* may be NULL when !registration); they are copied during the call. */
int ss_create(const ss_config *cfg, const float *mapLx, const float *mapLy, const float *mapRx,
              const float *mapRy, const float *a1, const float *a2, const float *a3,
              ss_engine **out);
/* Extension: calibration as MATRICES instead of seven H x W planes.  The reference generates the planes on
 * the host (python/py_package/sensor/simsense_component.py:177-215 rectification maps through
 * cv2.initUndistortRectifyMap, :308-325 registration planes) and reads them from HBM every frame; here
 * the kernels evaluate them per pixel from the 3x3 matrices (row-major, float64):
 *   registration  a(u,v) = reg_m * [u, v, 1]^T            (a1,a2,a3 = its three components, as float32)
 *   rectification [X,Y,W] = rect_inv_* * [u, v, 1]^T,  map = (fx * X/W + cx, fy * Y/W + cy)
 *                 with rect_inv_* = (P[:3,:3] * R)^-1 of that camera's stereoRectify result and
 *                 (fx, fy, cx, cy) the undistorted IR camera matrix (no lens distortion, as in the reference call).
 * No plane is uploaded, stored or read: -7 * rows * cols * 4 bytes of traffic per frame (26 MB at 1280x720). */
typedef struct {
  double reg_m[9];
  double rect_inv_left[9], rect_inv_right[9]; /* ignored when cfg->rectified */
  double ir_fx, ir_fy, ir_cx, ir_cy;
} ss_calibration;
int ss_create_calibrated(const ss_config *cfg, const ss_calibration *cal, ss_engine **out);
/* E:80 / C:484-541 */
int ss_destroy(ss_engine *e);

/* Region of interest, P:75-99 (bbox, bbox_start_x, bbox_start_y, bbox_width, bbox_height).
 * Validated here (inside the image, >=1 px); the reference validates nothing. */
typedef struct {
  int32_t enabled;
  uint32_t x, y, width, height;
} ss_bbox;

/* E:57-58 / C:308-330: host uint8 [batch][rows][cols] pairs.  Synchronous like the reference:
 * returns when the depth map is complete. */
int ss_compute_host_u8(ss_engine *e, const uint8_t *left, const uint8_t *right,
                       const ss_bbox *bbox);
/* E:60-61 / C:332-345: device float32 RGBA [batch][rows][cols][4]; the IR value is the R channel
 * (C:45-62).  Ordered after everything already enqueued on `stream` (a cudaStream_t; the special
 * handles cudaStreamLegacy (0x1) and cudaStreamPerThread (0x2) are accepted) and NOT synchronised:
 * `stream` is made to wait for the result, so the result pointers are valid for work enqueued on it
 * after this call.  stream = 0 (or the engine's own stream, ss_get_stream) means "the inputs are
 * already complete": no ordering is established, the work simply follows the engine's previous
 * frame.  The reference orders by a cudaDeviceSynchronize at the start of the frame (C:547); a
 * caller that wants that behaviour passes cudaStreamLegacy and calls ss_synchronize afterwards
 * (this is what the Python binding does by default).  The colour image of
 * ss_get_rgb_point_cloud_* is ordered after the legacy default stream. */
int ss_compute_device_rgba_f32(ss_engine *e, const void *left, const void *right,
                               const ss_bbox *bbox, void *stream);
/* The same for PITCHED views: environment n starts env_pitch_bytes * n after the base pointer, image row y
 * row_pitch_bytes * y after that; a pixel is always 4 consecutive floats.  This is how a window of
 * sapien's BatchedCamera buffer ([N,H,W,4] float32 CudaArrayHandle with byte strides,
 * src/sapien_renderer/batched_render_system.cpp:76-84) or any sliced tensor is consumed in place; order
 * the frame after the renderer with `stream` = the stream given to BatchedCamera::setCudaStream
 * (include/sapien/sapien_renderer/batched_render_system.h:35, the stream its external-semaphore wait
 * is enqueued on, batched_render_system.cpp:141-145). */
int ss_compute_device_rgba_f32_pitched(ss_engine *e, const void *left, const void *right,
                                       size_t env_pitch_bytes, size_t row_pitch_bytes,
                                       const ss_bbox *bbox, void *stream);
/* Extension: device uint8 [batch][rows][cols] pairs, same ordering rule. */
int ss_compute_device_u8(ss_engine *e, const void *left, const void *right, const ss_bbox *bbox,
                         void *stream);
/* Extension: ASYNCHRONOUS host-input frames.  ss_submit_host_u8 enqueues the uploads, the frame and the delivery of
 * its depth map into out_host (float32 [batch][out_rows][out_cols]; page-locked for the transfers to overlap; may be
 * NULL for "device result only") and returns at once with a ticket; ss_wait_frame blocks until that frame has been
 * delivered.  Up to TWO frames per lane are in flight: uploads and front-end of frame k+1 run while frame k aggregates, and the
 * read-back of frame k runs under frame k+1 (inputs, outputs and the upload buffers are double-buffered inside the
 * engine); a third submit first waits for the oldest frame.  left/right (and out_host) must stay valid until the
 * frame's ticket has been waited for.  ss_compute_host_u8 is the synchronous form of the same path. */
int ss_submit_host_u8(ss_engine *e, const uint8_t *left, const uint8_t *right, const ss_bbox *bbox,
                      float *out_host, size_t capacity_bytes, uint64_t *ticket);
int ss_wait_frame(ss_engine *e, uint64_t ticket);
/* Extension: page-locked host memory for the buffers of ss_submit_host_u8 / ss_bind_output_host, for bindings that have
 * no CUDA runtime of their own (the Python binding hands such buffers out as the arrays get_ndarray() returns, so that a
 * strict compute() + get_ndarray() caller gets its depth map without an extra host copy). */
int ss_alloc_host(size_t bytes, void **ptr);
int ss_free_host(void *ptr);
/* Extension: orders the NEXT compute of this engine after everything already enqueued on `stream`
 * (a producer stream other than the one passed to the compute call, e.g. the `stream` entry of a
 * __cuda_array_interface__ v3 input).  No host synchronisation. */
int ss_wait_stream(ss_engine *e, void *stream);
/* Waits for the last enqueued compute. */
int ss_synchronize(ss_engine *e);

/* E:71-72 */
int ss_get_output_shape(const ss_engine *e, uint32_t *rows, uint32_t *cols);
/* E:69-70 */
int ss_get_input_shape(const ss_engine *e, uint32_t *rows, uint32_t *cols);
/* E:66 / C:384-388 */
int ss_get_device(const ss_engine *e, int32_t *device);
/* Extension: the engine's PUBLIC stream (a cudaStream_t).  It runs no kernels; it completes, in submission order,
 * behind every frame: enqueue consumers of a frame's result on it (or record events on it around a run of frames).
 * Passing it as the `stream` of a compute call means "the inputs are complete, no ordering".  The reference keeps
 * its three streams private (core.h:95-97). */
int ss_get_stream(const ss_engine *e, void **stream);
/* Extension: number of lanes.  A lane is a complete set of streams and buffers; with several, consecutive frames
 * rotate over them and overlap on the GPU (the latency-bound head and tail of one frame's kernels are filled
 * by the other frames': C1 +15 % frames/s with two, +17 % with three), results in submission order on the public stream.  Batched engines
 * fill the machine by themselves and use one lane.  Borrowed result pointers stay valid until the next compute. */
int ss_get_lanes(const ss_engine *e, int32_t *lanes);

/* E:63 getMat2d / P:101: depth float32 [batch][out_rows][out_cols] copied to `out`. */
int ss_get_depth_host(ss_engine *e, float *out, size_t capacity_bytes);
/* Extension (no reference counterpart; the reference stages every read-back through an engine-owned
 * pageable buffer, C:347-362): bind a HOST buffer [batch][out_rows][out_cols] float32 -- pinned for
 * the copies to be asynchronous -- that every following compute streams its depth map into, band by
 * band, while the final aggregation pass is still running (the pass runs in column segments; the
 * finished columns are post-processed and copied behind each segment).  ss_get_depth_host() with
 * this same pointer then only waits for the frame.  NULL unbinds.  The buffer must stay valid until
 * it is unbound or the engine is destroyed.  Results are bit-identical to the unbound path. */
int ss_bind_output_host(ss_engine *e, float *out, size_t capacity_bytes);
/* E:67 getCudaPtr / P:103-111: borrowed device pointer. */
int ss_get_depth_device(ss_engine *e, void **ptr);
/* E:65 getPointCloudMat2d / P:122: float32 [batch][out_rows*out_cols][3]. */
int ss_get_point_cloud_host(ss_engine *e, float *out, size_t capacity_bytes);
/* E:68 getPointCloudCudaPtr / P:113-120 */
int ss_get_point_cloud_device(ss_engine *e, void **ptr);
/* E:69 getRgbPointCloudMat2d / P:124-128: rgba = device float32 [batch][out_rows][out_cols][4];
 * result float32 [batch][out_rows*out_cols][6] (x,y,z,r,g,b). */
int ss_get_rgb_point_cloud_host(ss_engine *e, const void *rgba_device, float *out,
                                size_t capacity_bytes);
/* E:70 getRgbPointCloudCudaPtr / P:130-139 */
int ss_get_rgb_point_cloud_device(ss_engine *e, const void *rgba_device, void **ptr);

/* Extension: the point cloud of the last frame WITHOUT a host synchronisation (the reference's device getters block,
 * C:409-411, 452): the kernel is enqueued behind the frame and *ptr is valid for work ordered on the engine's public
 * stream (ss_get_stream) after this call.  rgba_device = NULL: xyz [..][3]; else xyz + rgb [..][6]. */
int ss_enqueue_point_cloud(ss_engine *e, const void *rgba_device, void **ptr);

/* E:73-79 / C:457-482.  Validated against the ranges of
 * python/py_package/sensor/simsense_component.py:54-134; take effect at the next compute. */
int ss_set_ir_noise_parameters(ss_engine *e, float speckle_shape, float speckle_scale,
                               float gaussian_mu, float gaussian_sigma);
int ss_set_penalties(ss_engine *e, int32_t p1, int32_t p2);
int ss_set_census_window_size(ss_engine *e, int32_t width, int32_t height);
int ss_set_matching_block_size(ss_engine *e, int32_t width, int32_t height);
int ss_set_uniqueness_ratio(ss_engine *e, int32_t uniq_ratio);
int ss_set_lr_max_diff(ss_engine *e, int32_t lr_max_diff);

/* Parity/debug access to the intermediate buffers the reference keeps as protected members
 * (E:83-93).  Copies stage `name` of batch element `index` to host.  Names: "im0","im1" (u8,
 * images fed to census), "census0","census1" (u32), "cost","L0","L1","L2" (u16 volumes),
 * "L3","LAll" (u16 volumes, only with keep_stages), "disp_wta","disp_lr","disp_med" (f32),
 * "disp_right" (u16), "disp_full","depth" (f32, full IR size).  *bytes returns the size. */
int ss_get_stage_host(ss_engine *e, const char *name, int32_t index, void *out,
                      size_t capacity_bytes, size_t *bytes);

/* Per-stage device time of the last profiled compute (ms); enable with ss_set_profiling(e,1).
 * Replaces the compile-time PRINT_RUNTIME printf timing (config.h:26, C:549-780). */
int ss_set_profiling(ss_engine *e, int32_t enabled);
/* Sums over all computes since the previous call: ms[i] = total device time of stage names[i],
 * *frames = number of computes covered.  Synchronises the engine stream and resets the log. */
int ss_get_stage_times(ss_engine *e, const char **names, float *ms, int32_t capacity,
                       int32_t *count, int32_t *frames);
/* Number of kernels this library launches per compute call with the current settings. */
int ss_get_launches_per_compute(ss_engine *e, int32_t *count);

/* Device ordinal that owns a device pointer (replaces getCudaPtrDevice, python/pybind/sapien.cpp:291);
 * *device = -1 for host memory. */
int ss_pointer_device(const void *ptr, int32_t *device);

const char *ss_last_error(void);
const char *ss_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SS_B200_H */
