"""Host-side mirror of the reference's ``SimSenseComponent``
(python/py_package/sensor/simsense_component.py:25-338), re-hosted as a plain Python class
(no ``sapien.Component`` / ECS): same constructor arguments, same parameter validation and error
types, same forwarding methods; the engine behind it is the B200-native ``DepthSensorEngine``.

Extensions: ``device`` / ``batch`` / ``keep_stages`` keyword arguments, and ``on_add_to_scene`` may
be called with ``scene=None`` (``build()`` is an alias) because there is no scene graph here.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

from ..pose import Pose
from .calibration import calibrate

# The native module (sapien_b200.simsense: DepthSensorEngine, CudaArray) is imported when the engine is
# built, not at module import: the parameter validation, presets and calibration of this file are plain
# host logic that test infrastructure may use without mapping the CUDA library.


def _is_int(*vals) -> bool:
    return all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in vals)


def _odd_pair(a, b, limit) -> bool:
    return _is_int(a, b) and a > 0 and b > 0 and a % 2 == 1 and b % 2 == 1 and a * b <= limit


def validate_parameters(rgb_resolution, ir_resolution, ir_speckle_noise, ir_thermal_noise, census_width,
                        census_height, max_disp, block_width, block_height, p1_penalty, p2_penalty,
                        uniqueness_ratio, lr_max_diff, median_filter_size) -> None:
    """Raises TypeError exactly where the reference does (simsense_component.py:54-134)."""
    if not _is_int(rgb_resolution[0], rgb_resolution[1]):
        raise TypeError("RGB resolution (width and height) must be integer")
    if not _is_int(ir_resolution[0], ir_resolution[1]) or min(ir_resolution[0], ir_resolution[1]) < 32:
        raise TypeError("Infrared resolution (width and height) must be integer and no less than 32")
    if ir_speckle_noise > 0 and ir_thermal_noise <= 0:
        raise TypeError("ir_speckle_noise > 0, Infrared noise simulation is on. ir_thermal_noise must also be positive")
    if not _odd_pair(census_width, census_height, 65):
        raise TypeError("census_width and census_height must be positive odd integers and their product should be no larger than 65")
    if not _is_int(max_disp) or not 32 <= max_disp <= 1024:
        raise TypeError("max_disp must be integer and within range [32, 1024]")
    if not _odd_pair(block_width, block_height, 256):
        raise TypeError("block_width and block_height must be positive odd integers and their product should be no larger than 256")
    if not _is_int(p1_penalty, p2_penalty) or not 0 < p1_penalty < p2_penalty < 224:
        raise TypeError("p1_penalty must be positive integer less than p2_penalty and p2_penalty be positive integer less than 224")
    if not _is_int(uniqueness_ratio) or not 0 <= uniqueness_ratio <= 255:
        raise TypeError("uniqueness_ratio must be positive integer and no larger than 255")
    if not _is_int(lr_max_diff) or not -1 <= lr_max_diff <= 255:
        raise TypeError("lr_max_diff must be integer and within the range [0, 255]")
    if median_filter_size not in (1, 3, 5, 7):
        raise TypeError("Median filter size choices are 1, 3, 5, 7")


class SimSenseComponent:
    DEFAULT_SPECKLE_SHAPE = 1333.33
    DEFAULT_GAUSSIAN_SIGMA = 0.25
    DEFAULT_GAUSSIAN_MU = 0

    def __init__(self, rgb_resolution: tuple, ir_resolution: tuple, rgb_intrinsic: np.ndarray,
                 ir_intrinsic: np.ndarray, trans_pose_l: Pose, trans_pose_r: Pose, min_depth: float,
                 max_depth: float, ir_noise_seed: int, ir_speckle_noise: float, ir_thermal_noise: float,
                 rectified: bool, census_width: int, census_height: int, max_disp: int, block_width: int,
                 block_height: int, p1_penalty: int, p2_penalty: int, uniqueness_ratio: int,
                 lr_max_diff: int, median_filter_size: int, depth_dilation: bool, *, device: int = -1,
                 batch: int = 1, keep_stages: bool = False, device_calibration: bool = True):
        validate_parameters(rgb_resolution, ir_resolution, ir_speckle_noise, ir_thermal_noise, census_width,
                            census_height, max_disp, block_width, block_height, p1_penalty, p2_penalty,
                            uniqueness_ratio, lr_max_diff, median_filter_size)
        self.rgb_resolution = tuple(int(v) for v in rgb_resolution)
        self.ir_resolution = tuple(int(v) for v in ir_resolution)
        self.rgb_intrinsic = rgb_intrinsic
        self.ir_intrinsic = ir_intrinsic
        self.trans_pose_l = trans_pose_l
        self.trans_pose_r = trans_pose_r
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.ir_noise_seed = ir_noise_seed
        self.ir_speckle_noise = ir_speckle_noise
        self.ir_thermal_noise = ir_thermal_noise
        self.rectified = rectified
        self.census_width = census_width
        self.census_height = census_height
        self.max_disp = max_disp
        self.block_width = block_width
        self.block_height = block_height
        self.p1_penalty = p1_penalty
        self.p2_penalty = p2_penalty
        self.uniqueness_ratio = uniqueness_ratio
        self.lr_max_diff = lr_max_diff
        self.median_filter_size = median_filter_size
        self.depth_dilation = depth_dilation
        self.device = device
        self.batch = batch
        self.keep_stages = keep_stages
        # True: the engine gets the calibration as 3x3 matrices and evaluates the rectification maps and the
        # registration planes per pixel; False: the reference's seven H x W planes (simsense_component.py:177-215,308-325)
        self.device_calibration = device_calibration
        self._default_speckle_shape = self.DEFAULT_SPECKLE_SHAPE
        self._default_gaussian_sigma = self.DEFAULT_GAUSSIAN_SIGMA
        self._default_gaussian_mu = self.DEFAULT_GAUSSIAN_MU
        self._engine = None  # sapien_b200.simsense.DepthSensorEngine once added to a scene
        self.calibration = None

    # -- noise parameters as the engine wants them (simsense_component.py:169-175) --------------
    def noise_parameters(self, speckle: Optional[float] = None, thermal: Optional[float] = None):
        speckle = self.ir_speckle_noise if speckle is None else speckle
        thermal = self.ir_thermal_noise if thermal is None else thermal
        shape = 0.0 if speckle == 0 else self._default_speckle_shape / speckle
        return shape, speckle / self._default_speckle_shape, self._default_gaussian_mu, self._default_gaussian_sigma * thermal

    def on_add_to_scene(self, scene=None) -> None:
        cal = calibrate(self.ir_resolution, self.rgb_resolution, self.ir_intrinsic, self.rgb_intrinsic,
                        self.trans_pose_l, self.trans_pose_r, planes=not self.device_calibration)
        self.calibration = cal
        shape, scale, mu, sigma = self.noise_parameters()
        k = np.asarray(self.rgb_intrinsic, dtype=float)
        (iw, ih), (rw, rh) = self.ir_resolution, self.rgb_resolution
        from ..simsense import DepthSensorEngine  # fails loudly without the CUDA extension: there is no CPU fallback

        self._engine = DepthSensorEngine(
            ih, iw, rh, rw, cal.focal_len, cal.baseline_len, self.min_depth, self.max_depth,
            self.ir_noise_seed, shape, scale, mu, sigma, self.rectified, self.census_width,
            self.census_height, self.max_disp, self.block_width, self.block_height, self.p1_penalty,
            self.p2_penalty, self.uniqueness_ratio, self.lr_max_diff, self.median_filter_size,
            cal.map_lx, cal.map_ly, cal.map_rx, cal.map_ry, cal.a1, cal.a2, cal.a3,
            cal.b[0], cal.b[1], cal.b[2], self.depth_dilation, k[0][0], k[1][1], k[0][1], k[0][2], k[1][2],
            device=self.device, batch=self.batch, keep_stages=self.keep_stages,
            calibration=cal.matrices() if self.device_calibration else None)

    build = on_add_to_scene

    def on_remove_from_scene(self, scene=None) -> None:
        self._engine = None

    def _eng(self):
        if self._engine is None:
            raise RuntimeError("simsense component is not added to scene")
        return self._engine

    def compute(self, left: Union[np.ndarray, "CudaArray"], right: Union[np.ndarray, "CudaArray"],
                bbox_start: tuple = None, bbox_size: tuple = None) -> None:
        eng = self._eng()
        if bbox_start is not None and bbox_size is not None:
            eng.compute(left, right, True, *bbox_start, *bbox_size)
        else:
            eng.compute(left, right, False, 0, 0, 0, 0)

    def get_ndarray(self) -> np.ndarray:
        return self._eng().get_ndarray()

    def get_cuda(self) -> "CudaArray":
        return self._eng().get_cuda()

    def get_point_cloud_ndarray(self) -> np.ndarray:
        return self._eng().get_point_cloud_ndarray()

    def get_point_cloud_cuda(self) -> "CudaArray":
        return self._eng().get_point_cloud_cuda()

    def get_rgb_point_cloud_ndarray(self, rgba_cuda) -> np.ndarray:
        return self._eng().get_rgb_point_cloud_ndarray(rgba_cuda)

    def get_rgb_point_cloud_cuda(self, rgba_cuda) -> "CudaArray":
        return self._eng().get_rgb_point_cloud_cuda(rgba_cuda)
