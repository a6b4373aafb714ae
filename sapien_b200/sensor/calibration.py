"""Host-side calibration math for the stereo sensor: extrinsics from poses, OpenCV rectification,
registration planes.  Behavioural spec: python/py_package/sensor/simsense_component.py:166-215 and
:308-338 of the reference.  Runs once per sensor at set-up, never on the per-frame path."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..pose import Pose

# ROS (x forward, y left, z up) -> OpenCV (x right, y down, z forward); simsense_component.py:328-337
_ROS_TO_CV = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


def pose_to_cv_extrinsic(pose: Pose) -> np.ndarray:
    return _ROS_TO_CV @ np.linalg.inv(pose.to_transformation_matrix().astype(float))


def registration_planes(ir_size, k_ir: np.ndarray, k_rgb: np.ndarray, ir_to_rgb: np.ndarray):
    """a1,a2,a3 [h,w] float64 and b (3,) with [u',v',1]*z' = a(u,v)*z + b  (simsense_component.py:308-325)."""
    w, h = ir_size
    m = k_rgb @ ir_to_rgb[:3, :3] @ np.linalg.inv(k_ir)
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    pix = np.stack([u, v, np.ones_like(u)], axis=-1)
    a = np.einsum("ij,hwj->hwi", m, pix)
    b = (k_rgb @ ir_to_rgb[:3, 3:]).reshape(3)
    return a[..., 0], a[..., 1], a[..., 2], b


@dataclass
class Calibration:
    focal_len: float
    baseline_len: float
    map_lx: np.ndarray
    map_ly: np.ndarray
    map_rx: np.ndarray
    map_ry: np.ndarray
    a1: np.ndarray
    a2: np.ndarray
    a3: np.ndarray
    b: np.ndarray
    # the same calibration as 3x3 float64 matrices (DepthSensorEngine(calibration=...)): the engine then evaluates
    # the planes per pixel and none of the seven H x W arrays above is uploaded or read
    reg_m: np.ndarray = None          # K_rgb R_(ir->rgb) K_ir^-1: a(u,v) = reg_m [u,v,1]^T
    rect_inv_l: np.ndarray = None     # (P1[:3,:3] R1)^-1
    rect_inv_r: np.ndarray = None     # (P2[:3,:3] R2)^-1
    ir_camera: tuple = None           # (fx, fy, cx, cy) of the IR camera matrix

    def matrices(self):
        return (self.reg_m, self.rect_inv_l, self.rect_inv_r, self.ir_camera)


def calibrate(ir_size, rgb_size, k_ir, k_rgb, pose_l: Pose, pose_r: Pose, planes: bool = True) -> Calibration:
    """ir_size / rgb_size are (width, height).  Same OpenCV calls and arguments as the reference
    (stereoRectify alpha=1, initUndistortRectifyMap CV_32F), simsense_component.py:177-215."""
    import cv2

    k_ir = np.asarray(k_ir, dtype=float)
    k_rgb = np.asarray(k_rgb, dtype=float)
    rgb_pose = Pose()
    ex_rgb = pose_to_cv_extrinsic(rgb_pose)
    ex_l = pose_to_cv_extrinsic(rgb_pose * pose_l)
    ex_r = pose_to_cv_extrinsic(rgb_pose * pose_r)
    l2r = ex_r @ np.linalg.inv(ex_l)
    l2rgb = ex_rgb @ np.linalg.inv(ex_l)
    r1, r2, p1, p2, q, _, _ = cv2.stereoRectify(
        cameraMatrix1=k_ir, distCoeffs1=None, cameraMatrix2=k_ir, distCoeffs2=None,
        imageSize=tuple(ir_size), R=l2r[:3, :3], T=l2r[:3, 3:], alpha=1.0, newImageSize=tuple(ir_size))
    reg_m = k_rgb @ l2rgb[:3, :3] @ np.linalg.inv(k_ir)
    mats = dict(reg_m=reg_m, rect_inv_l=np.linalg.inv(p1[:3, :3] @ r1), rect_inv_r=np.linalg.inv(p2[:3, :3] @ r2),
                ir_camera=(float(k_ir[0][0]), float(k_ir[1][1]), float(k_ir[0][2]), float(k_ir[1][2])))
    if not planes:  # matrices only: no H x W array is generated (cv2 is used for the 3x3 stereoRectify alone)
        empty = np.zeros((0,), np.float32)
        b = (k_rgb @ l2rgb[:3, 3:]).reshape(3)
        return Calibration(float(q[2][3]), float(1.0 / q[3][2]), empty, empty, empty, empty, empty, empty, empty, b, **mats)
    map_lx, map_ly = cv2.initUndistortRectifyMap(k_ir, None, r1, p1, tuple(ir_size), cv2.CV_32F)
    map_rx, map_ry = cv2.initUndistortRectifyMap(k_ir, None, r2, p2, tuple(ir_size), cv2.CV_32F)
    a1, a2, a3, b = registration_planes(ir_size, k_ir, k_rgb, l2rgb)
    return Calibration(float(q[2][3]), float(1.0 / q[3][2]), map_lx, map_ly, map_rx, map_ry, a1, a2, a3, b, **mats)
