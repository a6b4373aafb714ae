// kernels.h -- host-callable launchers of the sm_100a kernels (one .cu per stage group).
// All launchers enqueue on `stream` and return the cudaError_t of the launch.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ssb {

// ---------------------------------------------------------------------------- front.cu
struct FrontParams {
  // sources: exactly one of src_u8 / src_rgba per image is non-null.  [N][frows][fcols](x4)
  const uint8_t *left_u8, *right_u8;
  const float *left_rgba, *right_rgba;
  // source pitches in ELEMENTS of the source type (u8 texels / floats): one image row, one environment.
  // Packed inputs: frows*fcols(*4) and fcols(*4).  A float RGBA pixel is always 4 consecutive floats.
  size_t src_env;
  uint32_t src_row;
  const float *mapLx, *mapLy, *mapRx, *mapRy; // null when rectified (or when the maps are evaluated from matrices)
  // matrix calibration (ss_create_calibrated): when cal_maps != 0 the rectification map of a pixel is evaluated
  // in double from rect_inv_* and the IR camera matrix instead of being read from the planes above
  int cal_maps;
  double rinvL[9], rinvR[9];
  double ir_fx, ir_fy, ir_cx, ir_cy;
  int frows, fcols;                           // full IR size
  int bx, by;                                 // ROI origin (0,0 when no bbox)
  int rows, cols;                             // matched (ROI) size
  int cw, ch;                                 // census window
  int N;
  uint8_t *im0, *im1;                         // [N][rows][cols] prepared images (census input)
  uint32_t *census0, *census1;                // [N][rows][cols]
  // optional IR noise (speckle_shape > 0): applied to the source texel before remap
  float speckle_shape, speckle_scale, gaussian_mu, gaussian_sigma;
  uint64_t seed;
  uint64_t frame; // frame counter (decorrelates successive frames)
  // optional: fill the registration canvas [canvas_n floats] with canvas_fill (initRgbDepth,
  // camera.cu:170-177) while the front-end runs, instead of a launch of its own
  float *canvas;
  size_t canvas_n;
  float canvas_fill;
  // -1: both images in one launch; 0 / 1: only the left / right image (7x7 census path), so that the
  // left image can be processed while the right one is still being uploaded
  int only_image;
};
cudaError_t launch_front(const FrontParams &p, cudaStream_t stream);

// ----------------------------------------------------------------------------- cost.cu
// C[n][y][x][d] = sum over the bw x bh replicate-border block of popc(cL(y,x) ^ cR(y,max(x-d,0)));
// bits = number of significant census bits (upper bound of one Hamming distance)
cudaError_t launch_cost(const uint32_t *cL, const uint32_t *cR, uint16_t *C, int N, int rows,
                        int cols, int D, int bw, int bh, int bits, cudaStream_t stream);

// ----------------------------------------------------------------------------- aggr.cu
// true when the packed-u16 DPX path is valid for this configuration
bool aggr_fast_supported(int D, int cmax, int P1, int P2);
// true when the two plain path volumes (L1, L2) may be stored 12-bit packed (1.5 D bytes per pixel)
bool aggr_pack12_supported(int D, int cmax, int P2);
struct AggrBuffers {
  const uint16_t *C; // cost volume
  uint16_t *L1;      // right->left path
  uint16_t *L2;      // top->bottom path
  uint16_t *S3;      // L1+L2+L3 (may alias L2 unless pack12)
  int pack12;        // L1 / L2 hold 12-bit packed pixels (fast path without stage materialisation only); S3 must not alias L2
  uint16_t *dbgL0, *dbgL3, *dbgLAll; // optional (keep_stages), else null
  float *dispL;      // [N][rows][cols] WTA left disparity (uniqueness + sub-pixel), -1 invalid
  uint16_t *dispR;   // [N][rows][cols] WTA right disparity
};
struct AggrMarks { // optional per-kernel timing marks (engine profiling mode)
  void (*mark)(void *ctx, const char *name);
  void *ctx;
};
// The 4 path aggregations + blend + winner-takes-all in two parts: the three plain passes (s_aux is a
// second stream that overlaps the two independent first passes; ev[0..1] are scratch events), then the
// final pass (left->right + blend + winner-takes-all).  With progress counters the final pass reports, per row, when the columns
// < seg_end[k] (multiples of 32, ascending, < cols) are finished: progress[k] reaches N*rows when
// every row is that far -- disparities of the columns < seg_end[k] - D are then final (the right
// disparity of a pixel completes D-1 columns later).  The caller zeroes the counters beforehand.
cudaError_t launch_aggr_passes(const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2,
                               int uniq, cudaStream_t stream, cudaStream_t s_aux, cudaEvent_t *ev,
                               const AggrMarks *marks = nullptr);
cudaError_t launch_aggr_final(const AggrBuffers &b, int N, int rows, int cols, int D, int P1, int P2,
                              int uniq, cudaStream_t stream, uint32_t *progress, int nseg, const int *seg_end);
// Generic (any D, 32-bit math) fallback with the same contract; needs scratch volumes.
cudaError_t launch_aggr_wta_generic(const AggrBuffers &b, uint16_t *L0scratch, int N, int rows,
                                    int cols, int D, int P1, int P2, int uniq,
                                    cudaStream_t stream);

// ----------------------------------------------------------------------------- post.cu
struct PostParams {
  int N;
  int rows, cols;   // matched size
  int frows, fcols; // full IR size
  int bx, by;
  int bbox;         // 1: ROI paste into a zeroed full-size map
  int lr_max_diff, mf_size;
  float focal, baseline, min_depth, max_depth;
  const float *dispL;      // [N][rows][cols] from WTA
  const uint16_t *dispR;   // [N][rows][cols]
  float *disp_lr;          // optional stage out [N][rows][cols]
  float *disp_med;         // [N][rows][cols]
  float *disp_full;        // [N][frows][fcols] (== disp_med layout when !bbox; may alias)
  float *depth;            // [N][frows][fcols]
  // registration
  int registration, dilation;
  const float *a1, *a2, *a3; // [frows][fcols]; all null: a(u,v) = reg_m * [u,v,1]^T evaluated per pixel
  double reg_m[9];
  float b1, b2, b3;
  int rgb_rows, rgb_cols;
  float *canvas;   // [N][rgb_rows][rgb_cols] splat target
  int canvas_prefilled; // 1: the front-end kernel already filled the canvas with max_depth
  float *out;      // final depth: [N][rgb_rows][rgb_cols] or [N][frows][fcols]
  // Column band (all zero = the whole frame).  LR check + median + depth + splat run on the matched
  // columns [xa, xb) (xa % 32 == 0), dilation + range clamp on the RGB columns [ua, ub)
  // (ua % 4 == 0); the bands of one frame are launched left to right on one stream and replace the
  // whole-frame call (the canvas is prefilled by the front-end).  No ROI mode.
  int xa, xb, ua, ub;
};
cudaError_t launch_post(const PostParams &p, cudaStream_t stream, int *launches);
// one-warp kernel that returns once *ctr >= target (the progress counters of launch_aggr_final)
cudaError_t launch_gate(const uint32_t *ctr, uint32_t target, cudaStream_t stream);

cudaError_t launch_point_cloud(const float *depth, const float *rgba, float *pc, int N, int rows,
                               int cols, float fx, float fy, float skew, float cx, float cy,
                               cudaStream_t stream);

} // namespace ssb
