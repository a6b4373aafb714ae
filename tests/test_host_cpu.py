"""Host-side logic that needs no GPU: the mirrored SimSenseComponent / StereoDepthSensorConfig
validation and presets (python/py_package/sensor/simsense_component.py:54-134,
stereodepth.py:31-202), calibration math, Pose, the synthetic input generator and the roofline
byte model."""
import hashlib

import numpy as np
import pytest

from oracle import configs
from sapien_b200 import synth
from sapien_b200.pose import Pose
from sapien_b200.sensor.calibration import calibrate, registration_planes
from sapien_b200.sensor.simsense_component import validate_parameters
from sapien_b200.sensor.stereodepth import StereoDepthSensorConfig

GOOD = dict(rgb_resolution=(1920, 1080), ir_resolution=(1280, 720), ir_speckle_noise=0.0, ir_thermal_noise=0.0,
            census_width=7, census_height=7, max_disp=128, block_width=7, block_height=7, p1_penalty=8,
            p2_penalty=32, uniqueness_ratio=15, lr_max_diff=1, median_filter_size=3)


def test_default_parameters_are_valid():
    validate_parameters(**GOOD)


@pytest.mark.parametrize("over", [
    dict(rgb_resolution=(1920.0, 1080)), dict(ir_resolution=(31, 720)), dict(ir_resolution=(1280, 31.5)),
    dict(ir_speckle_noise=1.0, ir_thermal_noise=0.0), dict(census_width=6), dict(census_width=0),
    dict(census_width=11, census_height=7), dict(max_disp=31), dict(max_disp=1025), dict(max_disp=64.0),
    dict(block_width=4), dict(block_width=17, block_height=17), dict(p1_penalty=0), dict(p1_penalty=32, p2_penalty=32),
    dict(p2_penalty=224), dict(uniqueness_ratio=-1), dict(uniqueness_ratio=256), dict(lr_max_diff=-2),
    dict(lr_max_diff=256), dict(median_filter_size=2), dict(median_filter_size=9),
])
def test_parameter_validation_raises_type_error_like_the_reference(over):
    with pytest.raises(TypeError):
        validate_parameters(**{**GOOD, **over})


@pytest.mark.parametrize("ok", [dict(census_width=13, census_height=5), dict(max_disp=32), dict(max_disp=1024),
                                dict(block_width=15, block_height=15), dict(p1_penalty=1, p2_penalty=223),
                                dict(uniqueness_ratio=255), dict(lr_max_diff=-1), dict(lr_max_diff=255),
                                dict(median_filter_size=7)])
def test_parameter_validation_accepts_the_range_edges(ok):
    validate_parameters(**{**GOOD, **ok})


def test_config_presets_and_defaults():
    c = StereoDepthSensorConfig("D415")
    assert c.rgb_resolution == (1920, 1080) and c.ir_resolution == (1280, 720)
    assert c.ir_intrinsic[0, 0] == 920.0 and c.rgb_intrinsic[0, 0] == 1380.0
    d = StereoDepthSensorConfig()
    assert d.ir_resolution == (848, 480)
    for cfg in (c, d):  # stereodepth.py:138-202
        assert (cfg.min_depth, cfg.max_depth, cfg.ir_noise_seed) == (0.2, 10.0, 0)
        assert (cfg.census_width, cfg.census_height, cfg.max_disp, cfg.block_width, cfg.block_height) == (7, 7, 128, 7, 7)
        assert (cfg.p1_penalty, cfg.p2_penalty, cfg.uniqueness_ratio, cfg.lr_max_diff, cfg.median_filter_size) == (8, 32, 15, 1, 3)
        assert cfg.rectified is True and cfg.depth_dilation is True
    with pytest.raises(ValueError):
        StereoDepthSensorConfig("D999")


def test_d415_calibration_matches_the_survey_probe():
    """f = 920 px, b = 54.5 mm, identity maps, A = diag(1.5,1.5,1), B = (24.15,0,0)."""
    c = StereoDepthSensorConfig("D415")
    cal = calibrate(c.ir_resolution, c.rgb_resolution, c.ir_intrinsic, c.rgb_intrinsic, c.trans_pose_l, c.trans_pose_r)
    assert abs(cal.focal_len - 920.0) < 1e-3
    assert abs(cal.baseline_len - 0.0545) < 1e-6
    xs, ys = np.meshgrid(np.arange(1280, dtype=np.float32), np.arange(720, dtype=np.float32))
    assert np.abs(cal.map_lx - xs).max() < 1e-2 and np.abs(cal.map_ly - ys).max() < 1e-2
    assert np.abs(cal.map_rx - xs).max() < 1e-2 and np.abs(cal.map_ry - ys).max() < 1e-2
    assert np.allclose(cal.a1, 1.5 * xs, atol=1e-3) and np.allclose(cal.a2, 1.5 * ys, atol=1e-3) and np.allclose(cal.a3, 1.0)
    assert np.allclose(cal.b, [24.15, 0.0, 0.0], atol=1e-4)


def test_d435_calibration_probe():
    c = StereoDepthSensorConfig("D435")
    cal = calibrate(c.ir_resolution, c.rgb_resolution, c.ir_intrinsic, c.rgb_intrinsic, c.trans_pose_l, c.trans_pose_r)
    assert abs(cal.focal_len - 430.14) < 0.01 and abs(cal.baseline_len - 0.050157) < 1e-5
    assert np.allclose(cal.b, [9.0813, 0.0440, 1.565e-4], atol=2e-3)


def test_registration_planes_identity():
    k = np.array([[100.0, 0, 50], [0, 100, 40], [0, 0, 1]])
    a1, a2, a3, b = registration_planes((8, 6), k, k, np.eye(4))
    u, v = np.meshgrid(np.arange(8), np.arange(6))
    assert np.allclose(a1, u) and np.allclose(a2, v) and np.allclose(a3, 1) and np.allclose(b, 0)


def test_pose_algebra():
    a = Pose([1, 2, 3], [np.cos(0.3), np.sin(0.3), 0, 0])
    b = Pose([-1, 0.5, 2], [np.cos(0.7), 0, np.sin(0.7), 0])
    ab = (a * b).to_transformation_matrix()
    assert np.allclose(ab, a.to_transformation_matrix() @ b.to_transformation_matrix(), atol=1e-5)
    assert np.allclose((a * a.inv()).to_transformation_matrix(), np.eye(4), atol=1e-5)
    assert np.allclose(Pose(a.to_transformation_matrix()).to_transformation_matrix(), a.to_transformation_matrix(), atol=1e-5)


def test_synthetic_pair_is_deterministic_and_consistent():
    l1, r1, d1 = synth.make_pair(64, 96, 32, seed=3)
    l2, r2, d2 = synth.make_pair(64, 96, 32, seed=3)
    assert np.array_equal(l1, l2) and np.array_equal(r1, r2) and np.array_equal(d1, d2)
    assert l1.dtype == np.uint8 and d1.min() >= 0 and d1.max() < 32
    ys, xs = np.mgrid[0:64, 0:96]
    ok = xs - d1 >= 0
    # a left pixel is found at x - d in the right image unless a nearer surface occludes it
    same = r1[ys[ok], (xs - d1)[ok]] == l1[ok]
    assert same.mean() > 0.85
    l3, _, _ = synth.make_pair(64, 96, 32, seed=4)
    assert not np.array_equal(l1, l3)


def test_to_rgba_round_trips_through_the_reference_conversion(oracle):
    img = np.arange(256, dtype=np.uint8).reshape(16, 16)
    rgba = synth.to_rgba(img)
    assert rgba.shape == (16, 16, 4) and rgba.dtype == np.float32
    assert np.array_equal(oracle.float2uint8(rgba), img)  # trunc(R*255), core.cu:45-62
    edge = np.zeros((1, 4, 4), np.float32)
    edge[0, :, 0] = [-0.5, 1.5, 0.999, 0.0039]
    assert oracle.float2uint8(edge[0][None])[0].tolist() == [0, 255, 254, 0]


def test_algorithmic_bytes_match_the_survey_table():
    """C1 2.519 GB, C3 0.842 GB, C4 93.6 MB, C5 10.92 GB per frame."""
    assert abs(configs.algorithmic_bytes(configs.params("C1"), rgba_input=True) / 1e9 - 2.519) < 0.002
    # the other rows of the survey table include optional stages (remap) our C3..C5 presets leave off: 2 %
    assert abs(configs.algorithmic_bytes(configs.params("C3"), rgba_input=True) / 0.842e9 - 1) < 0.02
    assert abs(configs.algorithmic_bytes(configs.params("C4"), rgba_input=True) / 93.6e6 - 1) < 0.02
    assert abs(configs.algorithmic_bytes(configs.params("C5"), rgba_input=True) / 10.92e9 - 1) < 0.02
    c2 = configs.algorithmic_bytes(configs.params("C2"), rgba_input=True, bbox=configs.BBOX_C2, point_cloud="xyzrgb")
    assert abs(c2 / 0.818e9 - 1) < 0.02


@pytest.mark.parametrize("model,roll", [("D415", 0.0), ("D435", 0.0), ("D415", 0.5)])
def test_matrix_calibration_reproduces_the_planes(model, roll):
    """Device-side calibration (ss_create_calibrated) evaluates per pixel what the reference tabulates on the host
    (simsense_component.py:177-215, 308-325).  Emulated here with the kernels' float64 operation order: the
    registration planes come out byte-identical as float32; the rectification maps agree with cv2's CV_32F planes to
    float32 rounding of values near zero and give the SAME nearest-neighbour source pixel everywhere (remap snaps
    every coordinate, camera.cu:83-119)."""
    import math

    c = StereoDepthSensorConfig(model)
    pose_r = c.trans_pose_r
    if roll:
        a = math.radians(roll) / 2
        pose_r = pose_r * Pose([0, 0, 0], [math.cos(a), math.sin(a), 0, 0])
    cal = calibrate(c.ir_resolution, c.rgb_resolution, c.ir_intrinsic, c.rgb_intrinsic, c.trans_pose_l, pose_r)
    only = calibrate(c.ir_resolution, c.rgb_resolution, c.ir_intrinsic, c.rgb_intrinsic, c.trans_pose_l, pose_r, planes=False)
    assert only.a1.size == 0 and only.map_lx.size == 0 and np.array_equal(only.reg_m, cal.reg_m) and np.array_equal(only.b, cal.b)
    w, h = c.ir_resolution
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    m = cal.reg_m
    for i, plane in enumerate((cal.a1, cal.a2, cal.a3)):
        mine = ((m[i, 0] * u + m[i, 1] * v) + m[i, 2]).astype(np.float32)
        assert np.array_equal(mine.view(np.uint32), plane.astype(np.float32).view(np.uint32)), f"a{i + 1}"
    fx, fy, cx, cy = cal.ir_camera
    for inv, mx, my in ((cal.rect_inv_l, cal.map_lx, cal.map_ly), (cal.rect_inv_r, cal.map_rx, cal.map_ry)):
        X = (u * inv[0, 0] + v * inv[0, 1]) + inv[0, 2]
        Y = (u * inv[1, 0] + v * inv[1, 1]) + inv[1, 2]
        W = (u * inv[2, 0] + v * inv[2, 1]) + inv[2, 2]
        iw = 1.0 / W
        gx, gy = (fx * (X * iw) + cx).astype(np.float32), (fy * (Y * iw) + cy).astype(np.float32)
        assert np.abs(gx - mx).max() < 1e-4 and np.abs(gy - my).max() < 1e-4
        assert (gx.view(np.uint32) != mx.view(np.uint32)).mean() < 2e-3  # (only the ~0 entries of column / row 0 differ)
        snap = lambda a, hi: np.clip(np.sign(a) * np.floor(np.abs(a) + 0.5), 0, hi)  # noqa: E731  roundf + clamp
        assert np.array_equal(snap(gx, w - 1), snap(mx, w - 1)) and np.array_equal(snap(gy, h - 1), snap(my, h - 1))
