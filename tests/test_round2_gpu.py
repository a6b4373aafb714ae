"""GPU parity tests added in round 2 (run with -m gpu on a B200): the production (keep_stages=False)
kernels at the headline size, batched device-RGBA input, the sensor facade, the raw C ABI driven
through ctypes, stream ordering of asynchronously produced inputs, the host-upload ordering of the
generic census path, and env sharding with real engines."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import configs
from sapien_b200 import synth
from tests.common import assert_depth_close, assert_stages_equal, get_stage, make_engine, variant

pytestmark = pytest.mark.gpu

from oracle import REF_SO, RefEngine  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libsimsense_ref.so not built")

FAST_STAGES = ("census0", "census1", "cost", "disp_wta", "disp_right", "disp_med", "depth")


def test_c1_production_kernels_vs_oracle(native, oracle):
    """The instantiation bench.py times at C1 -- DBG=false templates, the 5-rows-per-block pinned-role
    mapping of the final pass (720 rows / 148 SMs) and S3 aliasing L2 in place -- against the oracle:
    cost volume, both WTA disparities, median, depth and the final map."""
    prm = configs.params("C1")
    left, right = configs.pair(prm, seed=0)
    ref = oracle.pipeline(prm, left, right, volumes=True)
    eng = make_engine(native, prm)  # keep_stages=False
    eng.compute(left, right)
    assert_stages_equal(eng, prm, ref, names=FAST_STAGES)
    assert_depth_close(eng.get_ndarray(), ref["out"])
    # device RGBA input (what bench.py's `value` feeds) gives the same bits as host u8
    import torch

    host = eng.get_ndarray().copy()
    eng.compute(torch.from_numpy(synth.to_rgba(left)).cuda(), torch.from_numpy(synth.to_rgba(right)).cuda())
    assert np.array_equal(eng.get_ndarray().view(np.uint32), host.view(np.uint32))


@pytest.mark.parametrize("geom", [(160, 600, 128), (136, 740, 64), (200, 593, 96)])
def test_five_rows_per_block_geometries_vs_oracle(native, oracle, geom):
    """593..740 rows of one environment put 5 rows into every block of the final pass (pinned producer /
    consumer roles): production kernels vs the oracle on narrow images of that height."""
    cols, rows, d = geom
    prm = configs._sensor_params("D415", max_disp=d, rectified=False, roll_deg=0.5,
                                 scale=(cols, rows, (cols * 3) // 2, (rows * 3) // 2))
    left, right = configs.pair(prm, seed=rows)
    ref = oracle.pipeline(prm, left, right)
    eng = make_engine(native, prm)
    eng.compute(left, right)
    assert_stages_equal(eng, prm, ref, names=FAST_STAGES)
    assert_depth_close(eng.get_ndarray(), ref["out"])


@pytest.mark.parametrize("cfg,n", [("C3", 3), ("C4", 5)])
def test_batched_device_rgba_input_vs_oracle(native, oracle, cfg, n):
    """[N,H,W,4] float32 CUDA batches (the BatchedCamera layout, and what bench.py feeds C3/C4/C5):
    every environment of one batched call equals the oracle run on that environment alone."""
    import torch

    prm = configs.params(cfg)
    pairs = [configs.pair(prm, seed=200 + s) for s in range(n)]
    tl = torch.from_numpy(synth.to_rgba(np.stack([p[0] for p in pairs]))).cuda()
    tr = torch.from_numpy(synth.to_rgba(np.stack([p[1] for p in pairs]))).cuda()
    assert tuple(tl.shape) == (n, prm.rows, prm.cols, 4)
    eng = make_engine(native, prm, batch=n)
    eng.compute(tl, tr)
    out = eng.get_cuda().torch().cpu().numpy()
    assert out.shape == (n, prm.rgb_rows, prm.rgb_cols)
    for i, (l, r) in enumerate(pairs):
        ref = oracle.pipeline(prm, l, r, volumes=False)
        assert_stages_equal(eng, prm, ref, names=("im0", "im1", "census0", "census1", "disp_wta", "disp_right", "disp_med", "depth"), index=i)
        assert_depth_close(out[i], ref["out"], what=f"env {i}")


@pytest.mark.parametrize("cfg,n", [("small435odd", 9), ("small96", 13), ("small435", 8)])
def test_half_warp_final_pass_vs_oracle(native, oracle, cfg, n):
    """Batches in the throughput regime (more rows than one wave of the one-row final pass holds) at D = 64 / 96 run the
    final pass with two rows per warp pair (aggr_wta2_kernel): every environment against the oracle, including an odd
    total number of rows (the last pair has one row) and D = 96 (16 whole lanes of 6 disparities)."""
    prm = configs.params(cfg)
    assert n * prm.rows > 5 * 148
    pairs = [configs.pair(prm, seed=400 + s) for s in range(n)]
    eng = make_engine(native, prm, batch=n)
    for rep in range(2):  # twice: the second call reuses every ring / barrier
        eng.compute(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]))
    out = eng.get_ndarray()
    for i, (l, r) in enumerate(pairs):
        ref = oracle.pipeline(prm, l, r, volumes=False)
        assert_stages_equal(eng, prm, ref, names=("disp_wta", "disp_right", "disp_med", "depth"), index=i)
        assert_depth_close(out[i], ref["out"], what=f"env {i}")


def test_strided_batched_camera_view(native, oracle):
    """BatchedCamera hands out [N,H,W,4] views with byte strides (batched_render_system.cpp:55-100): a view into a
    larger allocation (row pitch and env pitch larger than the packed ones) must give the same result as the packed
    copy."""
    import torch

    prm = configs.params("small435")
    n = 3
    pairs = [configs.pair(prm, seed=300 + s) for s in range(n)]
    packed_l = torch.from_numpy(synth.to_rgba(np.stack([p[0] for p in pairs]))).cuda()
    packed_r = torch.from_numpy(synth.to_rgba(np.stack([p[1] for p in pairs]))).cuda()
    big_l = torch.full((n + 1, prm.rows + 3, prm.cols + 5, 4), -1.0, device="cuda")
    big_r = torch.full((n + 1, prm.rows + 3, prm.cols + 5, 4), -1.0, device="cuda")
    view_l = big_l[1:, 2:2 + prm.rows, 4:4 + prm.cols]
    view_r = big_r[1:, 2:2 + prm.rows, 4:4 + prm.cols]
    view_l.copy_(packed_l)
    view_r.copy_(packed_r)
    assert not view_l.is_contiguous()
    eng = make_engine(native, prm, batch=n)
    eng.compute(packed_l, packed_r)
    want = eng.get_ndarray().copy()
    eng.compute(view_l, view_r)
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want.view(np.uint32))
    ref = oracle.pipeline(prm, *pairs[1], volumes=False)
    assert_depth_close(want[1], ref["out"])
    # a channel stride other than 4 bytes is not an image the renderer can produce: rejected, not misread
    with pytest.raises(RuntimeError):
        eng.compute(big_l[1:, :prm.rows, :prm.cols].transpose(2, 3)[..., :4], packed_r)


def test_sensor_facade_on_gpu(native, oracle):
    """StereoDepthSensor(StereoDepthSensorConfig("D415")) -> set_pictures -> compute_depth(bbox) -> get_depth /
    get_pointcloud(with_rgb=True), against the oracle (python/py_package/sensor/stereodepth.py:294-475)."""
    import torch

    from sapien_b200.sensor import StereoDepthSensor, StereoDepthSensorConfig

    cfg = StereoDepthSensorConfig("D415")
    cfg.ir_speckle_noise = 0.0  # bit-exact parity: noise off (statistical parity is tested separately)
    cfg.ir_thermal_noise = 0.0
    cfg.max_disp = 64
    sensor = StereoDepthSensor(cfg)
    prm = configs._sensor_params("D415", max_disp=64, rectified=True)
    left, right = configs.pair(prm, seed=4)
    rgba = synth.make_rgb(prm.rgb_rows, prm.rgb_cols, 1)
    rgba_t = torch.from_numpy(rgba).cuda()
    with pytest.raises(RuntimeError):
        sensor.take_picture()
    sensor.set_pictures(torch.from_numpy(synth.to_rgba(left)).cuda(), torch.from_numpy(synth.to_rgba(right)).cuda(), rgba_t)
    sensor.compute_depth()
    ref = oracle.pipeline(prm, left, right, volumes=False)
    assert_depth_close(sensor.get_depth(), ref["out"])
    d_cuda = sensor.get_depth_cuda()
    assert tuple(d_cuda.shape) == (1080, 1920) and d_cuda.typestr == "f4"
    assert np.array_equal(d_cuda.torch().cpu().numpy().view(np.uint32), sensor.get_depth().view(np.uint32))
    pc = sensor.get_pointcloud(with_rgb=True)
    want = oracle.pointcloud(sensor.get_depth(), rgba, prm.main_fx, prm.main_fy, prm.main_skew, prm.main_cx, prm.main_cy)
    np.testing.assert_allclose(pc, want, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sensor.get_pointcloud_cuda().torch().cpu().numpy(), want[:, :3], rtol=1e-4, atol=1e-6)
    # ROI compute (manualtest/stereodepth_bbox.py:118-119) with uint8 host pictures
    sensor.set_pictures(left, right)
    sensor.compute_depth(bbox_start=(100, 100), bbox_size=(640, 360))
    ref_roi = oracle.pipeline(prm, left, right, bbox=(100, 100, 640, 360), volumes=False)
    assert_depth_close(sensor.get_depth(), ref_roi["out"])
    # runtime setters go through to the engine with the reference's validation
    sensor.set_penalties(4, 60)
    sensor.set_uniqueness_ratio(30)
    sensor.compute_depth()
    ref2 = oracle.pipeline(variant(prm, p1=4, p2=60, uniq_ratio=30), left, right, volumes=False)
    assert_depth_close(sensor.get_depth(), ref2["out"])
    with pytest.raises(TypeError):
        sensor.set_penalties(60, 4)
    assert sensor.get_config().p1_penalty == 4
    # the STOCK configuration (D435, IR noise on: speckle 1.0, thermal 1.0) runs and stays close to the noise-free result
    stock = StereoDepthSensor(StereoDepthSensorConfig())
    p435 = configs.params("C3")
    l4, r4 = configs.pair(variant(p435, max_disp=128), seed=6)
    stock.set_pictures(l4, r4)
    stock.compute_depth()
    noisy = stock.get_depth()
    stock.set_ir_noise(0.0, 0.0)
    stock.compute_depth()
    clean = stock.get_depth()
    assert noisy.shape == clean.shape == (480, 848)
    both = (noisy > 0) & (clean > 0)
    assert both.mean() > 0.5 * (clean > 0).mean() and np.median(np.abs(noisy[both] - clean[both]) / clean[both]) < 0.02


def test_raw_c_abi_compute_through_ctypes(native, oracle):
    """The drop-in boundary itself: ss_create / ss_compute_host_u8 / ss_get_depth_host / ss_get_stage_host called
    through ctypes with plain pointers (no pybind), against the oracle."""
    from tests.test_cabi_cpu import LIB, SsConfig

    lib = C.CDLL(LIB)
    lib.ss_last_error.restype = C.c_char_p
    prm = configs.params("small435")
    left, right = configs.pair(prm, seed=21)
    pl = prm.planes()
    cfg = SsConfig(rows=prm.rows, cols=prm.cols, rgb_rows=prm.rgb_rows, rgb_cols=prm.rgb_cols, focal_len=prm.focal_len,
                   baseline_len=prm.baseline_len, min_depth=prm.min_depth, max_depth=prm.max_depth, ir_noise_seed=0,
                   rectified=int(prm.rectified), census_width=prm.census_width, census_height=prm.census_height,
                   max_disp=prm.max_disp, bf_width=prm.bf_width, bf_height=prm.bf_height, p1=prm.p1, p2=prm.p2,
                   uniq_ratio=prm.uniq_ratio, lr_max_diff=prm.lr_max_diff, mf_size=prm.mf_size, b1=prm.b1, b2=prm.b2,
                   b3=prm.b3, dilation=int(prm.dilation), main_fx=prm.main_fx, main_fy=prm.main_fy,
                   main_skew=prm.main_skew, main_cx=prm.main_cx, main_cy=prm.main_cy, registration=1, device=-1,
                   batch=1, keep_stages=0)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    eng = C.c_void_p()
    rc = lib.ss_create(C.byref(cfg), None, None, None, None, fp(pl["a1"]), fp(pl["a2"]), fp(pl["a3"]), C.byref(eng))
    assert rc == 0, lib.ss_last_error()
    try:
        out = np.empty((prm.rgb_rows, prm.rgb_cols), np.float32)
        assert lib.ss_get_depth_host(eng, fp(out), C.c_size_t(out.nbytes)) == 2  # SS_ERR_NOT_COMPUTED
        assert lib.ss_compute_host_u8(eng, fp(left), fp(right), None) == 0, lib.ss_last_error()
        assert lib.ss_get_depth_host(eng, fp(out), C.c_size_t(out.nbytes)) == 0, lib.ss_last_error()
        ref = oracle.pipeline(prm, left, right, volumes=False)
        assert_depth_close(out, ref["out"])
        dr = np.empty((prm.rows, prm.cols), np.uint16)
        nbytes = C.c_size_t()
        assert lib.ss_get_stage_host(eng, b"disp_right", 0, fp(dr), C.c_size_t(dr.nbytes), C.byref(nbytes)) == 0
        assert nbytes.value == dr.nbytes and np.array_equal(dr, ref["disp_right"])
        rows, cols = C.c_uint32(), C.c_uint32()
        assert lib.ss_get_output_shape(eng, C.byref(rows), C.byref(cols)) == 0
        assert (rows.value, cols.value) == (prm.rgb_rows, prm.rgb_cols)
        small = np.empty(16, np.float32)
        assert lib.ss_get_depth_host(eng, fp(small), C.c_size_t(small.nbytes)) == 1  # SS_ERR_INVALID: buffer too small
    finally:
        assert lib.ss_destroy(eng) == 0


def test_inputs_produced_asynchronously_on_the_default_stream(native):
    """compute(left_cuda, right_cuda) with the default stream=None must order the frame after work still queued
    on the caller's (legacy default) stream -- the reference does a device-wide sync at the start of the frame
    (core.cu:547).  The inputs are written by kernels queued BEHIND a long spin kernel."""
    import torch

    prm = configs.params("small435")
    left, right = configs.pair(prm, seed=31)
    eng = make_engine(native, prm)
    eng.compute(left, right)
    want = eng.get_ndarray().copy()
    l8, r8 = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    tl = torch.zeros((prm.rows, prm.cols, 4), device="cuda")
    tr = torch.zeros((prm.rows, prm.cols, 4), device="cuda")
    torch.cuda.synchronize()
    for rep in range(3):
        tl.zero_()
        tr.zero_()
        torch.cuda._sleep(40_000_000)  # ~20 ms on the default stream
        tl.copy_(((l8.float() + 0.5) / 255.0)[..., None].expand(-1, -1, 4))
        tr.copy_(((r8.float() + 0.5) / 255.0)[..., None].expand(-1, -1, 4))
        eng.compute(tl, tr, sync=False)       # host returns at once; the frame must wait for the copies above
        got = eng.get_cuda().torch().clone()  # ... and the default stream for the frame (clone runs on it)
        assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32)), f"rep {rep}"
    # a side stream announced through `stream=`
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        tl.zero_()
        torch.cuda._sleep(40_000_000)
        tl.copy_(((l8.float() + 0.5) / 255.0)[..., None].expand(-1, -1, 4))
        eng.compute(tl, tr, stream=side.cuda_stream, sync=False)
        got = eng.get_cuda().torch().clone()
    side.synchronize()
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("census", [(5, 5), (9, 7)])
def test_pinned_host_upload_is_ordered_before_the_generic_front_end(native, census):
    """Host u8 compute with a census window other than 7x7: the uploads go to engine-owned buffers on the main stream
    and the front-end runs on the helper stream -- it must wait for the DMA (pinned source, > 1 MB per image, so the
    copy is truly asynchronous).  Compared with the same images given as device arrays."""
    import torch

    prm = variant(configs.params("C1"), census_width=census[0], census_height=census[1])
    n = 2
    pairs = [configs.pair(prm, seed=50 + i) for i in range(2 * n)]
    eng = make_engine(native, prm, batch=n)
    for k in range(2):
        l = np.stack([pairs[2 * k + i][0] for i in range(n)])
        r = np.stack([pairs[2 * k + i][1] for i in range(n)])
        eng.compute(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
        want = eng.get_ndarray().copy()
        pl, pr = torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()
        # poison the engine's upload buffers with the OTHER pair first, so stale data cannot pass
        o = 1 - k
        eng.compute(np.stack([pairs[2 * o + i][0] for i in range(n)]), np.stack([pairs[2 * o + i][1] for i in range(n)]))
        eng.compute(pl.numpy(), pr.numpy())
        assert np.array_equal(eng.get_ndarray().view(np.uint32), want.view(np.uint32)), f"pair set {k}"


def test_sharded_stereo_depth_real_engines(native, oracle):
    """ShardedStereoDepth with real engines in one process: a shard of exactly ONE environment keeps its leading
    dimension, and pipelines=2 (two engines on their own streams) equals the single batched engine."""
    import torch

    from sapien_b200.sharding import ShardedStereoDepth

    prm = configs.params("C4")
    n = 5
    pairs = [configs.pair(prm, seed=400 + s) for s in range(n)]
    tl = torch.from_numpy(synth.to_rgba(np.stack([p[0] for p in pairs]))).cuda()
    tr = torch.from_numpy(synth.to_rgba(np.stack([p[1] for p in pairs]))).cuda()
    one = ShardedStereoDepth(prm.engine_args(), n_envs=1)
    d1 = one.compute(tl[:1], tr[:1])
    assert tuple(d1.shape) == (1, prm.rgb_rows, prm.rgb_cols)
    ref0 = oracle.pipeline(prm, *pairs[0], volumes=False)
    assert_depth_close(d1[0].cpu().numpy(), ref0["out"])
    assert one.gather_depth(d1) is d1  # single process: identity
    whole = ShardedStereoDepth(prm.engine_args(), n_envs=n)
    piped = ShardedStereoDepth(prm.engine_args(), n_envs=n, pipelines=2)
    assert [b - a for a, b in piped.blocks] == [3, 2]
    a = whole.compute(tl, tr).clone()
    b = piped.compute(tl, tr)
    assert tuple(b.shape) == (n, prm.rgb_rows, prm.rgb_cols)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    # rank 3 of 4 owns one environment of 5, rank 0 two: the blocks tile the batch
    parts = []
    for r in range(4):
        sh = ShardedStereoDepth(prm.engine_args(), n_envs=n, rank=r, world=4)
        parts.append(sh.compute(tl[sh.start:sh.stop], tr[sh.start:sh.stop]).clone())
    assert torch.equal(torch.cat(parts).view(torch.int32), a.view(torch.int32))


def _ks(a_u8: np.ndarray, b_u8: np.ndarray) -> float:
    """Two-sample Kolmogorov-Smirnov statistic of two uint8 samples (from their histograms)."""
    ca = np.cumsum(np.bincount(a_u8.ravel(), minlength=256)) / a_u8.size
    cb = np.cumsum(np.bincount(b_u8.ravel(), minlength=256)) / b_u8.size
    return float(np.abs(ca - cb).max())


@needs_ref
def test_ir_noise_statistical_parity_vs_reference(native):
    """simInfraredNoise (camera.cu:21-75): speckle Gamma(shape, scale) * I + thermal N(mu, sigma), rounded and clamped.
    The reference draws from per-pixel XORWOW states (48 B/pixel read and written back per image and frame); this
    engine from a stateless Philox stream keyed by (seed, texel, frame).  Same distribution, different numbers:
    compared on > 10^6 pixels per image -- per-intensity mean and variance, and a two-sample KS test on the noisy u8
    histograms -- for constant images and for a textured one, stock parameters (simsense_component.py:160-175) and a
    strong-noise setting (shape < 1 exercises the boost branch)."""
    n = 1024
    base = configs._sensor_params("D415", max_disp=32, rectified=True, scale=(n, n, n, n))
    ks_crit = 1.95 * np.sqrt(2.0 / (n * n))  # alpha = 0.001
    rng = np.random.Generator(np.random.PCG64(5))
    tex = rng.integers(0, 256, size=(n, n), dtype=np.uint8)
    flat = [np.full((n, n), v, np.uint8) for v in (0, 30, 120, 220, 255)]
    for shape, scale, mu, sigma in ((1333.33, 1 / 1333.33, 0.0, 0.25), (40.0, 1 / 40.0, 1.5, 3.0), (0.7, 1 / 0.7, 0.0, 1.0)):
        prm = variant(base, speckle_shape=shape, speckle_scale=scale, gaussian_mu=mu, gaussian_sigma=sigma, ir_noise_seed=1234)
        ours = make_engine(native, prm)
        ref = RefEngine(prm)
        for img in flat + [tex]:
            ours.compute(img, img)
            ref.compute_host(img, img)
            for side in (0, 1):
                a = get_stage(ours, prm, f"im{side}")
                b = ref.stage(f"noisyim{side}")
                d = _ks(a, b)
                assert d < ks_crit, f"shape={shape} image={int(img[0, 0]) if img is not tex else 'tex'} side={side}: KS {d:.5f} >= {ks_crit:.5f}"
                if img is tex:  # per-intensity moments on the textured image
                    src = img.ravel()
                    cnt = np.bincount(src, minlength=256).astype(np.float64)
                    fa, fb = a.ravel().astype(np.float64), b.ravel().astype(np.float64)
                    ma = np.bincount(src, weights=fa, minlength=256) / cnt
                    mb = np.bincount(src, weights=fb, minlength=256) / cnt
                    va = np.bincount(src, weights=fa * fa, minlength=256) / cnt - ma * ma
                    vb = np.bincount(src, weights=fb * fb, minlength=256) / cnt - mb * mb
                    se = np.sqrt((va + vb) / cnt) + 1e-9  # standard error of the difference of the means
                    assert np.all(np.abs(ma - mb) < 5.0 * se + 0.02), f"per-intensity means differ: max z {np.max(np.abs(ma - mb) / se):.2f}"
                    # ~4 100 pixels per intensity: a variance estimate has relative standard error ~sqrt(2/n)
                    tol = 6.0 * np.sqrt(2.0 / cnt) * np.sqrt(2.0) * np.maximum(va, vb) + 0.05
                    assert np.all(np.abs(va - vb) < tol), f"per-intensity variances differ: worst {np.max(np.abs(va - vb) / tol):.2f} x tolerance"
                    # pooled over all intensities the two variances agree much more tightly
                    assert abs(va.mean() - vb.mean()) < 0.01 * max(va.mean(), vb.mean()) + 0.01
                else:
                    fa, fb = a.astype(np.float64), b.astype(np.float64)
                    se = np.sqrt((fa.var() + fb.var()) / a.size) + 1e-9
                    assert abs(fa.mean() - fb.mean()) < 5.0 * se + 1e-3
                    assert abs(fa.var() - fb.var()) < 0.02 * max(fa.var(), fb.var()) + 0.01
        # the left and the right image use different streams, and successive frames too
        ours.compute(tex, tex)
        f1 = get_stage(ours, prm, "im0").copy()
        ours.compute(tex, tex)
        assert not np.array_equal(f1, get_stage(ours, prm, "im0")) and not np.array_equal(f1, get_stage(ours, prm, "im1"))
        ref.close()


@pytest.mark.parametrize("cfg,bbox", [("C1", None), ("small435", None), ("small", (8, 4, 64, 40)), ("C4", None)])
def test_matrix_calibration_equals_planes(native, cfg, bbox):
    """ss_create_calibrated: rectification maps and registration planes evaluated per pixel from 3x3 matrices (no
    H x W plane uploaded or read) give bit-identical stages and depth to the plane-fed engine -- C1 exercises the
    non-trivial (0.5 degree roll) remap, small435 / C4 the non-diagonal D435 registration."""
    import math

    from sapien_b200.pose import Pose
    from sapien_b200.sensor.calibration import calibrate
    from sapien_b200.sensor.stereodepth import StereoDepthSensorConfig

    prm = configs.params(cfg)
    model, roll, scale = {"C1": ("D415", 0.5, None), "small435": ("D435", 0.0, (128, 96, 128, 96)),
                          "small": ("D415", 0.5, (96, 64, 144, 96)), "C4": ("D435", 0.0, (256, 256, 256, 256))}[cfg]
    c = StereoDepthSensorConfig(model)
    k_ir, k_rgb = c.ir_intrinsic.copy(), c.rgb_intrinsic.copy()
    ir_size, rgb_size = c.ir_resolution, c.rgb_resolution
    if scale is not None:  # same rescaling as oracle/configs.py
        k_ir[0] *= scale[0] / ir_size[0]
        k_ir[1] *= scale[1] / ir_size[1]
        k_rgb[0] *= scale[2] / rgb_size[0]
        k_rgb[1] *= scale[3] / rgb_size[1]
        ir_size, rgb_size = (scale[0], scale[1]), (scale[2], scale[3])
    pose_r = c.trans_pose_r
    if roll:
        a = math.radians(roll) / 2
        pose_r = pose_r * Pose([0, 0, 0], [math.cos(a), math.sin(a), 0, 0])
    cal = calibrate(ir_size, rgb_size, k_ir, k_rgb, c.trans_pose_l, pose_r, planes=False)
    assert math.isclose(cal.focal_len, prm.focal_len) and math.isclose(cal.baseline_len, prm.baseline_len)
    args = list(prm.engine_args())
    empty = np.zeros((0,), np.float32)
    for i in range(24, 31):  # the seven plane arguments
        args[i] = empty
    planes = make_engine(native, prm, keep_stages=True)
    mats = native.DepthSensorEngine(*args, keep_stages=True, calibration=cal.matrices())
    left, right = configs.pair(prm, seed=77)
    bb = () if bbox is None else (True, *bbox)
    planes.compute(left, right, *bb)
    mats.compute(left, right, *bb)
    for st in ("im0", "im1", "census0", "census1", "disp_wta", "disp_right", "disp_med", "depth"):
        a, b = get_stage(mats, prm, st, bbox), get_stage(planes, prm, st, bbox)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), st
    assert np.array_equal(mats.get_ndarray().view(np.uint32), planes.get_ndarray().view(np.uint32))
    # banded host output works on the matrix path too
    import torch

    out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    fast = native.DepthSensorEngine(*args, calibration=cal.matrices())
    fast.bind_output(out)
    fast.compute(left, right)
    fast.get_ndarray(out=out)
    planes.compute(left, right)
    assert np.array_equal(out.view(np.uint32), planes.get_ndarray().view(np.uint32))


@pytest.mark.parametrize("cfg", ["C1", "C4", "small"])
def test_pipelined_host_frames_submit_wait(native, cfg):
    """ss_submit_host_u8 / ss_wait_frame: two frames in flight -- uploads and front-end of frame k+1 under frame k's
    aggregation, read-back of frame k under frame k+1.  Every delivered map must equal the same pair computed alone,
    also when synchronous, device-input, ROI and generic-census frames are mixed into the sequence."""
    import torch

    prm = configs.params(cfg)
    n = 7
    pairs = [configs.pair(prm, seed=500 + i) for i in range(n)]
    solo = make_engine(native, prm)
    want = []
    for l, r in pairs:
        solo.compute(l, r)
        want.append(solo.get_ndarray())
    pin = [(torch.from_numpy(l).pin_memory().numpy(), torch.from_numpy(r).pin_memory().numpy()) for l, r in pairs]
    outs = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    eng = make_engine(native, prm)
    for rep in range(3):
        tickets = []
        for i in range(n):
            if i >= 2:  # the output buffer of frame i-2 is about to be reused: consume it first
                eng.wait(tickets[i - 2])
                assert np.array_equal(outs[i % 2].view(np.uint32), want[i - 2].view(np.uint32)), f"frame {i - 2} (rep {rep})"
                outs[i % 2][:] = -5.0
            tickets.append(eng.submit(pin[i][0], pin[i][1], out=outs[i % 2]))
        for i in (n - 2, n - 1):
            eng.wait(tickets[i])
            assert np.array_equal(outs[i % 2].view(np.uint32), want[i].view(np.uint32)), f"frame {i} (rep {rep})"
        eng.wait(tickets[0])  # waiting again for an old ticket is harmless
    # getters after an un-waited asynchronous frame
    t = eng.submit(pin[3][0], pin[3][1], out=outs[0])
    assert np.array_equal(eng.get_cuda().torch().cpu().numpy().view(np.uint32), want[3].view(np.uint32))
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[3].view(np.uint32))
    eng.wait(t)
    assert np.array_equal(outs[0].view(np.uint32), want[3].view(np.uint32))
    # mixed sequence: async, sync host, device, async without host output, async ROI (no column bands), sync again
    dl, dr = torch.from_numpy(pairs[5][0]).cuda(), torch.from_numpy(pairs[5][1]).cuda()
    t0 = eng.submit(pin[0][0], pin[0][1], out=outs[0])
    eng.compute(*pairs[1])
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[1].view(np.uint32))
    eng.wait(t0)
    assert np.array_equal(outs[0].view(np.uint32), want[0].view(np.uint32))
    t2 = eng.submit(pin[2][0], pin[2][1], out=outs[1])
    eng.compute(dl, dr)
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[5].view(np.uint32))
    eng.wait(t2)
    assert np.array_equal(outs[1].view(np.uint32), want[2].view(np.uint32))
    t3 = eng.submit(pin[4][0], pin[4][1])  # no host delivery
    eng.wait(t3)
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[4].view(np.uint32))
    bbox = (16, 8, prm.cols // 2, prm.rows // 2)
    solo.compute(*pairs[6], True, *bbox)
    roi = solo.get_ndarray()
    t4 = eng.submit(pin[6][0], pin[6][1], outs[0], True, *bbox)
    t5 = eng.submit(pin[0][0], pin[0][1], out=outs[1])
    eng.wait(t4)
    assert np.array_equal(outs[0].view(np.uint32), roi.view(np.uint32))
    eng.wait(t5)
    assert np.array_equal(outs[1].view(np.uint32), want[0].view(np.uint32))
    with pytest.raises(TypeError):
        eng.wait(10 ** 9)


def test_pipelined_host_frames_generic_census(native):
    """Asynchronous host frames with a census window other than 7x7 (uploads on the main stream, generic front-end)."""
    import torch

    prm = variant(configs.params("small435"), census_width=9, census_height=7)
    pairs = [configs.pair(prm, seed=600 + i) for i in range(4)]
    solo = make_engine(native, prm)
    eng = make_engine(native, prm)
    outs = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    tickets = [None] * 4
    for i, (l, r) in enumerate(pairs):
        if i >= 2:
            eng.wait(tickets[i - 2])
            solo.compute(*pairs[i - 2])
            assert np.array_equal(outs[i % 2].view(np.uint32), solo.get_ndarray().view(np.uint32))
        tickets[i] = eng.submit(l, r, out=outs[i % 2])
    for i in (2, 3):
        eng.wait(tickets[i])
        solo.compute(*pairs[i])
        assert np.array_equal(outs[i % 2].view(np.uint32), solo.get_ndarray().view(np.uint32))


@pytest.mark.parametrize("over", [dict(max_disp=64), dict(max_disp=96), dict(max_disp=128), dict(max_disp=256),
                                  dict(max_disp=64, p1=20, p2=59), dict(max_disp=128, p1=30, p2=60)])
def test_packed_path_volumes_vs_oracle(native, oracle, over):
    """Production engines store the two plain path volumes (L1, L2) 12-bit packed when cmax + P2 < 4096 (storage only).
    L1 read back through ss_get_stage_host (unpacked on the host) and everything downstream must equal the oracle --
    one / two / four registers per lane, the partial-lane D = 96 case, and penalties just below (p2 = 59: 1176 + 2891 =
    4067) and just above (p2 = 60: 4116, packing off) the 12-bit limit."""
    prm = variant(configs._sensor_params("D415", max_disp=64, rectified=True, scale=(320, 120, 320, 120)), **over)
    left, right = configs.pair(prm, seed=prm.max_disp + prm.p2)
    ref = oracle.pipeline(prm, left, right)
    eng = make_engine(native, prm)
    eng.compute(left, right)
    assert np.array_equal(get_stage(eng, prm, "L1"), ref["L1"])
    assert_stages_equal(eng, prm, ref, names=FAST_STAGES)
    assert_depth_close(eng.get_ndarray(), ref["out"])
    batch = make_engine(native, prm, batch=3)
    batch.compute(np.stack([left, right, left]), np.stack([right, left, right]))
    assert np.array_equal(get_stage(batch, prm, "L1", index=2), ref["L1"])
    assert np.array_equal(batch.get_ndarray()[0].view(np.uint32), eng.get_ndarray().view(np.uint32))


@pytest.mark.parametrize("lanes", [2, 3])
def test_lanes_are_bit_identical_to_one_lane(native, lanes):
    """Frames enqueued back to back rotate over the engine's lanes (independent stream / buffer sets); every frame must
    equal what a one-lane engine computes for the same pair, for device inputs read back later and in any order."""
    import torch

    prm = configs.params("small128")
    pairs = [configs.pair(prm, seed=900 + i) for i in range(7)]
    solo = make_engine(native, prm, lanes=1)
    eng = make_engine(native, prm, lanes=lanes)
    assert eng.lanes == lanes and solo.lanes == 1
    want = []
    for l, r in pairs:
        solo.compute(l, r)
        want.append(solo.get_ndarray().copy())
    dev = [(torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()) for l, r in pairs]
    torch.cuda.synchronize()
    es = torch.cuda.ExternalStream(eng.cuda_stream)
    got = []
    for tl, tr in dev:  # no host synchronisation between the frames: the results are cloned in stream order
        eng.compute(tl, tr, stream=eng.cuda_stream, sync=False)
        with torch.cuda.stream(es):
            got.append(eng.get_cuda().torch().clone())
    es.synchronize()
    for i, g in enumerate(got):
        assert np.array_equal(g.cpu().numpy().view(np.uint32), want[i].view(np.uint32)), f"frame {i} differs with {lanes} lanes"


def test_strict_signature_results_are_never_overwritten(native):
    """compute(l, r) + get_ndarray() with nothing else: the map is delivered into a page-locked pool array which
    get_ndarray() hands out.  Arrays a caller keeps must stay intact however many frames follow (the pool only reuses
    arrays nobody references; beyond four kept results it falls back to fresh arrays), repeated get_ndarray() calls
    return equal data in distinct arrays, and writing into a returned array does not leak into later frames."""
    prm = configs.params("small435")
    pairs = [configs.pair(prm, seed=700 + i) for i in range(8)]
    solo = make_engine(native, prm, lanes=1)
    want = []
    for l, r in pairs:
        solo.compute(l, r)
        want.append(solo.get_ndarray(out=np.empty((prm.rgb_rows, prm.rgb_cols), np.float32)).copy())
    eng = make_engine(native, prm)
    kept = []
    for i, (l, r) in enumerate(pairs):  # keep every result: more than the pool holds
        eng.compute(l, r)
        a = eng.get_ndarray()
        b = eng.get_ndarray()
        assert a is not b and np.array_equal(a.view(np.uint32), b.view(np.uint32))
        kept.append(a)
        for j, k in enumerate(kept):
            assert np.array_equal(k.view(np.uint32), want[j].view(np.uint32)), f"result {j} changed after frame {i}"
    kept.clear()
    for rep in range(3):  # results dropped at once: the pool arrays are reused
        for i, (l, r) in enumerate(pairs):
            eng.compute(l, r)
            a = eng.get_ndarray()
            assert np.array_equal(a.view(np.uint32), want[i].view(np.uint32))
            a[:] = -1.0  # a caller may scribble over its own result
    # ROI frame and device frame in between
    eng.compute(*pairs[0], True, 8, 4, 64, 40)
    roi = eng.get_ndarray()
    solo.compute(*pairs[0], True, 8, 4, 64, 40)
    assert np.array_equal(roi.view(np.uint32), solo.get_ndarray().view(np.uint32))
    import torch

    eng.compute(torch.from_numpy(pairs[3][0]).cuda(), torch.from_numpy(pairs[3][1]).cuda())
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[3].view(np.uint32))


def test_stream_ordered_point_clouds(native, oracle):
    """get_[rgb_]point_cloud_cuda(sync=False): the kernel is enqueued behind the frame on its lane, nothing blocks the
    host; consumers ordered on the engine's public stream see the finished cloud -- frames back to back on two lanes."""
    import torch

    prm = configs.params("small435")
    pairs = [configs.pair(prm, seed=800 + i) for i in range(4)]
    rgba = synth.make_rgb(prm.rgb_rows, prm.rgb_cols, 2)
    rgba_t = torch.from_numpy(rgba).cuda()
    torch.cuda.synchronize()
    eng = make_engine(native, prm)
    es = torch.cuda.ExternalStream(eng.cuda_stream)
    dl = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in pairs]
    torch.cuda.synchronize()
    snaps, depths = [], []
    for i in range(4):
        eng.compute(dl[i][0], dl[i][1], stream=eng.cuda_stream, sync=False)
        pc = eng.get_rgb_point_cloud_cuda(rgba_t, sync=False) if i % 2 == 0 else eng.get_point_cloud_cuda(sync=False)
        with torch.cuda.stream(es):
            snaps.append(pc.torch().clone())
            depths.append(eng.get_cuda().torch().clone())
    es.synchronize()
    for i in range(4):
        want = oracle.pointcloud(depths[i].cpu().numpy(), rgba if i % 2 == 0 else None, prm.main_fx, prm.main_fy, prm.main_skew,
                                 prm.main_cx, prm.main_cy)
        np.testing.assert_allclose(snaps[i].cpu().numpy(), want, rtol=1e-4, atol=1e-6)
        ref = oracle.pipeline(prm, *pairs[i], volumes=False)
        assert_depth_close(depths[i].cpu().numpy(), ref["out"])


def test_invalid_pixels_that_splat_inside_the_image(native, oracle):
    """b3 > 0 with (b1/b3, b2/b3) inside the RGB image: every invalid disparity (z = 0) splats zRgb = b3 onto that one
    pixel (camera.cu:187-195 does not skip z = 0).  The banded host delivery assumes an RGB column is final once the
    matched columns that can reach it are done -- not true for that pixel -- so it must switch itself off: bound
    output, plain read-back and the oracle agree."""
    import torch

    prm = variant(configs.params("small435"), b1=60.0, b2=40.0, b3=1.0)
    left, right = configs.pair(prm, seed=900)
    ref = oracle.pipeline(prm, left, right, volumes=False)
    assert ref["out"][40, 60] == np.float32(1.0)  # the pixel all invalid disparities land on
    plain = make_engine(native, prm)
    plain.compute(left, right)
    assert_depth_close(plain.get_ndarray(), ref["out"])
    band = make_engine(native, prm)
    out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    band.bind_output(out)
    for _ in range(3):
        out[:] = -1.0
        band.compute(left, right)
        band.get_ndarray(out=out)
        assert np.array_equal(out.view(np.uint32), plain.get_ndarray().view(np.uint32))


@pytest.mark.parametrize("cfg,batch,wave", [("C4", 5, 2), ("small435", 7, 3), ("small", 3, 1)])
def test_batches_larger_than_one_wave(native, oracle, cfg, batch, wave):
    """A batch that does not fit the memory budget runs in waves of `wave` environments (front-end, cost volume and the
    four passes per wave on the same volumes, post-processing once for the whole batch).  On a 180 GB GPU that never
    happens by itself: a debug hook caps the wave size.  Device-RGBA, host and pipelined host inputs."""
    import ctypes
    import torch

    from tests.test_cabi_cpu import LIB

    lib = ctypes.CDLL(LIB)
    prm = configs.params(cfg)
    pairs = [configs.pair(prm, seed=950 + i) for i in range(batch)]
    l, r = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    whole = make_engine(native, prm, batch=batch)
    whole.compute(l, r)
    want = whole.get_ndarray().copy()
    ref = oracle.pipeline(prm, *pairs[batch - 1], volumes=False)
    assert_depth_close(want[batch - 1], ref["out"])
    lib.ssb_debug_set_max_wave(wave)
    try:
        eng = make_engine(native, prm, batch=batch)
    finally:
        lib.ssb_debug_set_max_wave(0)
    for _ in range(2):
        eng.compute(l, r)
        assert np.array_equal(eng.get_ndarray().view(np.uint32), want.view(np.uint32))
        eng.compute(torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda())
        assert np.array_equal(eng.get_cuda().torch().cpu().numpy().view(np.uint32), want.view(np.uint32))
    out = torch.empty((batch, prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    t = eng.submit(l, r, out=out)
    eng.wait(t)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    assert_stages_equal(eng, prm, ref, names=("census0", "disp_wta", "disp_right", "disp_med", "depth"), index=batch - 1)
