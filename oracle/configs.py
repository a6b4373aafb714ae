"""The BASELINE.json workloads as `Params` + synthetic inputs.
Test / bench infrastructure only."""
from __future__ import annotations

import functools
import math

import numpy as np

from . import Params


def _sensor_params(model: str, *, max_disp: int, rectified: bool, roll_deg: float = 0.0, scale=None, **over) -> Params:
    from sapien_b200.pose import Pose
    from sapien_b200.sensor.calibration import calibrate
    from sapien_b200.sensor.stereodepth import StereoDepthSensorConfig

    c = StereoDepthSensorConfig(model)
    ir_size, rgb_size = c.ir_resolution, c.rgb_resolution
    k_ir, k_rgb = c.ir_intrinsic.copy(), c.rgb_intrinsic.copy()
    if scale is not None:  # (ir_w, ir_h, rgb_w, rgb_h): rescale the intrinsics with the resolution
        sx, sy = scale[0] / ir_size[0], scale[1] / ir_size[1]
        k_ir[0] *= sx
        k_ir[1] *= sy
        rx, ry = scale[2] / rgb_size[0], scale[3] / rgb_size[1]
        k_rgb[0] *= rx
        k_rgb[1] *= ry
        ir_size, rgb_size = (scale[0], scale[1]), (scale[2], scale[3])
    pose_r = c.trans_pose_r
    if roll_deg:
        a = math.radians(roll_deg) / 2
        pose_r = pose_r * Pose([0, 0, 0], [math.cos(a), math.sin(a), 0, 0])
    cal = calibrate(ir_size, rgb_size, k_ir, k_rgb, c.trans_pose_l, pose_r)
    p = Params(
        rows=ir_size[1], cols=ir_size[0], rgb_rows=rgb_size[1], rgb_cols=rgb_size[0],
        focal_len=cal.focal_len, baseline_len=cal.baseline_len, min_depth=c.min_depth, max_depth=c.max_depth,
        rectified=rectified, census_width=c.census_width, census_height=c.census_height, max_disp=max_disp,
        bf_width=c.block_width, bf_height=c.block_height, p1=c.p1_penalty, p2=c.p2_penalty,
        uniq_ratio=c.uniqueness_ratio, lr_max_diff=c.lr_max_diff, mf_size=c.median_filter_size,
        map_lx=cal.map_lx, map_ly=cal.map_ly, map_rx=cal.map_rx, map_ry=cal.map_ry,
        a1=cal.a1.astype(np.float32), a2=cal.a2.astype(np.float32), a3=cal.a3.astype(np.float32),
        b1=float(cal.b[0]), b2=float(cal.b[1]), b3=float(cal.b[2]), dilation=c.depth_dilation,
        main_fx=float(k_rgb[0][0]), main_fy=float(k_rgb[1][1]), main_skew=float(k_rgb[0][1]),
        main_cx=float(k_rgb[0][2]), main_cy=float(k_rgb[1][2]))
    for k, v in over.items():
        setattr(p, k, v)
    return p


@functools.lru_cache(maxsize=None)
def params(name: str) -> Params:
    """C1..C5 of BASELINE.json `configs` (+ small variants used by the parity tests)."""
    if name == "C1":  # 1280x720 D415, D=128, remap exercised through a 0.5 deg roll of the right camera
        return _sensor_params("D415", max_disp=128, rectified=False, roll_deg=0.5)
    if name == "C1r":  # same, already rectified (stock config)
        return _sensor_params("D415", max_disp=128, rectified=True)
    if name == "C2":  # C1 + bbox (100,100) 640x360 + RGB point cloud
        return _sensor_params("D415", max_disp=128, rectified=False, roll_deg=0.5)
    if name == "C3":  # 848x480 D435, D=96
        return _sensor_params("D435", max_disp=96, rectified=True)
    if name == "C4":  # 256x256 low-res, D=64 (D435 intrinsics rescaled)
        return _sensor_params("D435", max_disp=64, rectified=True, scale=(256, 256, 256, 256))
    if name == "C5":  # 1920x1080, D=256 (D415 rescaled)
        return _sensor_params("D415", max_disp=256, rectified=True, scale=(1920, 1080, 1920, 1080))
    if name == "small":  # 96x64, D=32: oracle-in-milliseconds parity case with maps + registration
        return _sensor_params("D415", max_disp=32, rectified=False, roll_deg=0.5, scale=(96, 64, 144, 96))
    if name == "small435":  # 128x96, D=64, D435 (non-trivial a1..a3), rectified
        return _sensor_params("D435", max_disp=64, rectified=True, scale=(128, 96, 128, 96))
    if name == "small128":  # 160x96, D=128: the 32-columns-per-thread cost kernel and the 2-register SGM lanes at a sanitizer-friendly size
        return _sensor_params("D415", max_disp=128, rectified=False, roll_deg=0.5, scale=(160, 96, 240, 144))
    if name == "small435odd":  # 128x95, D=64: an odd number of rows (x an odd batch: the last row pair of the half-warp final pass is half empty)
        return _sensor_params("D435", max_disp=64, rectified=True, scale=(128, 95, 128, 95))
    if name == "small96":  # 144x64, D=96 (D435): the partial-lane SGM kernels and the half-warp final pass of batches
        return _sensor_params("D435", max_disp=96, rectified=True, scale=(144, 64, 144, 64))
    if name == "small256":  # 288x64, D=256
        return _sensor_params("D415", max_disp=256, rectified=True, scale=(288, 64, 288, 64))
    raise KeyError(name)


BBOX_C2 = (100, 100, 640, 360)  # x, y, w, h  (manualtest/stereodepth_bbox.py:118-119)


def pair(prm: Params, seed: int = 0):
    from sapien_b200.synth import make_pair

    left, right, _ = make_pair(prm.rows, prm.cols, prm.max_disp, seed)
    return left, right


def algorithmic_bytes(prm: Params, *, rgba_input: bool, bbox=None, point_cloud: str = "") -> int:
    """Normative per-frame byte count of SURVEY.md section 8(d) / BASELINE.md section 3."""
    s0 = prm.rows * prm.cols
    s = s0 if bbox is None else bbox[2] * bbox[3]
    r = prm.rgb_rows * prm.rgb_cols
    v = 2 * s * prm.max_disp
    total = 10 * s + (8 * s + v) + 8 * v + (v + 6 * s) + 8 * s0 + (20 * r + 24 * s0)
    if rgba_input:
        total += 34 * s0
    if not prm.rectified:
        total += 20 * s0
    if bbox is not None:
        total += 4 * s + 8 * s + 4 * s0
    if prm.lr_max_diff != 255:
        total += 10 * s
    if prm.mf_size != 1:
        total += 8 * s
    if point_cloud == "xyz":
        total += 16 * r
    elif point_cloud == "xyzrgb":
        total += 44 * r
    return int(total)
