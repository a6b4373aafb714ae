"""``StereoDepthSensorConfig`` / ``StereoDepthSensor``-compatible facade, minus the renderer.

Behavioural spec: python/py_package/sensor/stereodepth.py:31-202 (config presets and defaults)
and :205-475 (sensor).  The reference sensor owns three render cameras and an active light and
pulls its IR/RGB pictures from the Vulkan renderer (``take_picture``); the renderer is out of scope
here, so pictures are *given* to the sensor: ``set_pictures(left, right, rgba)`` takes
what ``RenderCameraComponent.get_picture_cuda("Color")`` would return (float32 RGBA CUDA arrays), or
uint8 numpy IR images.  Everything downstream (``compute_depth``, getters, runtime setters) keeps
the reference's names, arguments and error behaviour.  ``batch=N`` adds a leading env dimension.
"""
from __future__ import annotations

from copy import copy
from typing import Optional

import numpy as np

from ..pose import Pose
from .simsense_component import SimSenseComponent, _is_int, _odd_pair


class StereoDepthSensorConfig:
    """Same attributes, presets and defaults as the reference config (stereodepth.py:31-202)."""

    SUPPORTED_MODELS = ("D415", "D435")

    def __init__(self, model: str = "D435"):
        if model == "D415":  # stereodepth.py:42-78
            self.light_pattern = None
            self.rgb_resolution = (1920, 1080)
            self.ir_resolution = (1280, 720)
            self.rgb_intrinsic = np.array([[1380.0, 0.0, 960.0], [0.0, 1380.0, 540.0], [0.0, 0.0, 1.0]])
            self.ir_intrinsic = np.array([[920.0, 0.0, 640.0], [0.0, 920.0, 360.0], [0.0, 0.0, 1.0]])
            self.trans_pose_l = Pose([0, -0.0175, 0])
            self.trans_pose_r = Pose([0, -0.0720, 0])
            self.pose_rgb_irproj = Pose()
            self.active_light_fov = 1.57
        elif model == "D435":  # stereodepth.py:79-136
            self.light_pattern = None
            self.rgb_resolution = (848, 480)
            self.ir_resolution = (848, 480)
            self.rgb_intrinsic = np.array(
                [[605.12158203125, 0.0, 424.5927734375], [0.0, 604.905517578125, 236.668975830078], [0.0, 0.0, 1.0]])
            self.ir_intrinsic = np.array(
                [[430.139801025391, 0.0, 425.162841796875], [0.0, 430.139801025391, 235.276519775391], [0.0, 0.0, 1.0]])
            rot = [[9.99948919e-01, 9.67109110e-03, 2.94523709e-03],
                   [-9.70994867e-03, 9.99862015e-01, 1.34780351e-02],
                   [-2.81448336e-03, -1.35059441e-02, 9.99904811e-01]]

            def _pose(t):
                m = np.eye(4)
                m[:3, :3] = rot
                m[:3, 3] = t
                return Pose(m)

            self.trans_pose_l = _pose([1.56505470e-04, -1.48976548e-02, -1.15314942e-05])
            self.trans_pose_r = _pose([-3.28569353e-04, -6.50479272e-02, 6.65888772e-04])
            self.pose_rgb_irproj = _pose([0.0, -0.015 - 0.029, 0.0])
            self.active_light_fov = np.deg2rad(102.0)
        else:
            raise ValueError(f"Invalid sensor model, must be one of: {self.SUPPORTED_MODELS}")
        # algorithm defaults, stereodepth.py:138-202
        self.ir_camera_exposure = 0.01
        self.min_depth = 0.2
        self.max_depth = 10.0
        self.ir_noise_seed = 0
        self.ir_speckle_noise = 1.0
        self.ir_thermal_noise = 1.0
        self.rectified = True
        self.census_width = 7
        self.census_height = 7
        self.max_disp = 128
        self.block_width = 7
        self.block_height = 7
        self.p1_penalty = 8
        self.p2_penalty = 32
        self.uniqueness_ratio = 15
        self.lr_max_diff = 1
        self.median_filter_size = 3
        self.depth_dilation = True


class StereoDepthSensor:
    def __init__(self, config: StereoDepthSensorConfig, mount=None, pose: Pose = Pose(), *,
                 device: int = -1, batch: int = 1, keep_stages: bool = False, device_calibration: bool = True):
        self._config = config
        self._mount = mount
        self._pose = pose
        c = config
        self._ss = SimSenseComponent(
            c.rgb_resolution, c.ir_resolution, c.rgb_intrinsic, c.ir_intrinsic, c.trans_pose_l,
            c.trans_pose_r, c.min_depth, c.max_depth, c.ir_noise_seed, c.ir_speckle_noise,
            c.ir_thermal_noise, c.rectified, c.census_width, c.census_height, c.max_disp, c.block_width,
            c.block_height, c.p1_penalty, c.p2_penalty, c.uniqueness_ratio, c.lr_max_diff,
            c.median_filter_size, c.depth_dilation, device=device, batch=batch, keep_stages=keep_stages,
            device_calibration=device_calibration)
        self._ss.on_add_to_scene(None)
        self._left = self._right = self._rgba = None

    # ---- picture hand-over (replaces take_picture, stereodepth.py:275-292) ---------------------
    def set_pictures(self, left, right, rgba=None) -> None:
        self._left, self._right = left, right
        if rgba is not None:
            self._rgba = rgba

    def take_picture(self, infrared_only: bool = False):
        raise RuntimeError("Cannot take picture: the renderer is out of scope of sapien_b200; use set_pictures(left, right, rgba)")

    def compute_depth(self, bbox_start: tuple = None, bbox_size: tuple = None, left=None, right=None):
        if left is not None:
            self.set_pictures(left, right)
        if self._left is None or self._right is None:
            raise RuntimeError("no pictures: call set_pictures(left, right) first")
        self._ss.compute(self._left, self._right, bbox_start, bbox_size)

    # ---- runtime setters (stereodepth.py:309-417), same checks and messages --------------------
    def set_ir_noise(self, ir_speckle_noise: float, ir_thermal_noise: float):
        if ir_speckle_noise > 0 and ir_thermal_noise <= 0:
            raise TypeError("ir_speckle_noise > 0, Infrared noise simulation is on. ir_thermal_noise must also be positive")
        self._config.ir_speckle_noise = ir_speckle_noise
        self._config.ir_thermal_noise = ir_thermal_noise
        # the reference reads self._default_speckle_shape from the wrong object here (:321)
        self._ss._engine.set_ir_noise_parameters(*self._ss.noise_parameters(ir_speckle_noise, ir_thermal_noise))

    def set_census_window_size(self, census_width: int, census_height: int):
        if not _odd_pair(census_width, census_height, 65):
            raise TypeError("census_width and census_height must be positive odd integers and their product should be no larger than 65")
        self._config.census_width, self._config.census_height = census_width, census_height
        self._ss._engine.set_census_window_size(census_width, census_height)

    def set_matching_block_size(self, block_width: int, block_height: int):
        if not _odd_pair(block_width, block_height, 256):
            raise TypeError("block_width and block_height must be positive odd integers and their product should be no larger than 256")
        self._config.block_width, self._config.block_height = block_width, block_height
        self._ss._engine.set_matching_block_size(block_width, block_height)

    def set_penalties(self, p1_penalty: int, p2_penalty: int):
        if not _is_int(p1_penalty, p2_penalty) or not 0 < p1_penalty < p2_penalty < 224:
            raise TypeError("p1_penalty must be positive integer less than p2_penalty and p2_penalty be positive integer less than 224")
        self._config.p1_penalty, self._config.p2_penalty = p1_penalty, p2_penalty
        self._ss._engine.set_penalties(p1_penalty, p2_penalty)

    def set_uniqueness_ratio(self, uniqueness_ratio: int):
        if not _is_int(uniqueness_ratio) or not 0 <= uniqueness_ratio <= 255:
            raise TypeError("uniqueness_ratio must be positive integer and no larger than 255")
        self._config.uniqueness_ratio = uniqueness_ratio
        self._ss._engine.set_uniqueness_ratio(uniqueness_ratio)

    def set_lr_max_diff(self, lr_max_diff: int):
        if not _is_int(lr_max_diff) or not -1 <= lr_max_diff <= 255:
            raise TypeError("lr_max_diff must be integer within the range [0, 255]")
        self._config.lr_max_diff = lr_max_diff
        self._ss._engine.set_lr_max_diff(lr_max_diff)

    # ---- getters (stereodepth.py:419-475) --------------------------------------------------------
    def get_config(self):
        return copy(self._config)

    def get_pose(self):
        return copy(self._pose if self._mount is None else self._mount.get_pose() * self._pose)

    def get_rgba_cuda(self):
        if self._rgba is None:
            raise RuntimeError("no RGB picture: pass rgba to set_pictures()")
        return self._rgba

    def get_depth(self) -> np.ndarray:
        return self._ss.get_ndarray()

    def get_depth_cuda(self):
        return self._ss.get_cuda()

    def get_pointcloud(self, with_rgb: bool = False) -> np.ndarray:
        if with_rgb:
            return self._ss.get_rgb_point_cloud_ndarray(self.get_rgba_cuda())
        return self._ss.get_point_cloud_ndarray()

    def get_pointcloud_cuda(self, with_rgb: bool = False):
        if with_rgb:
            return self._ss.get_rgb_point_cloud_cuda(self.get_rgba_cuda())
        return self._ss.get_point_cloud_cuda()
