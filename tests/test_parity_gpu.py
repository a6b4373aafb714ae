"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, driven through the pybind layer
over the C ABI, against (a) the scalar C oracle and (b) the unmodified reference simsense CUDA code
(oracle/_ref).  Integer stages must match bit-for-bit; float outputs within the north_star
tolerances (<=1e-3 px sub-pixel disparity, <=1e-4 relative depth) -- in practice disparity and the
IR-frame depth are bit-exact too."""
import os

import numpy as np
import pytest

from oracle import REF_SO, RefEngine, configs
from sapien_b200 import synth
from tests import golden_util
from tests.common import (assert_depth_close, assert_stages_equal, get_stage, make_engine, variant)

pytestmark = pytest.mark.gpu

SMALL_VARIANTS = {
    "default": {},
    "rectified": dict(rectified=True),
    "no_lr": dict(lr_max_diff=255),
    "lr0": dict(lr_max_diff=0),
    "lr3": dict(lr_max_diff=3),
    "mf1": dict(mf_size=1),
    "mf5": dict(mf_size=5),
    "mf7": dict(mf_size=7),
    "bf1": dict(bf_width=1, bf_height=1),
    "bf3": dict(bf_width=3, bf_height=3),
    "bf5": dict(bf_width=5, bf_height=5),
    "bf9": dict(bf_width=9, bf_height=9),
    "bf3x5_generic_cost": dict(bf_width=3, bf_height=5),
    "bf11x1_generic_cost": dict(bf_width=11, bf_height=1),
    "census5": dict(census_width=5, census_height=5),
    "census9x7": dict(census_width=9, census_height=7),
    "census3": dict(census_width=3, census_height=3),
    "census13x5": dict(census_width=13, census_height=5),
    "uniq0": dict(uniq_ratio=0),
    "uniq50": dict(uniq_ratio=50),
    "uniq100": dict(uniq_ratio=100),
    "uniq180": dict(uniq_ratio=180),
    "p_small": dict(p1=1, p2=2),
    "p_large": dict(p1=100, p2=223),
    "no_dilation": dict(dilation=False),
    "d48": dict(max_disp=48),
    "d64": dict(max_disp=64),
    "d96": dict(max_disp=96),
    "d100": dict(max_disp=100),
    "d128_wider_than_image": dict(max_disp=128),
    "d160": dict(max_disp=160),
    "d256": dict(max_disp=256),
    "d33_generic_aggr": dict(max_disp=33),
    "d130_generic_aggr": dict(max_disp=130),
    "wide_regime_generic_aggr": dict(bf_width=15, bf_height=15, p1=100, p2=223),
    "depth_range": dict(min_depth=0.5, max_depth=1.0),
}


def run_ours(native, prm, left, right, bbox=None, **kw):
    eng = make_engine(native, prm, **kw)
    if bbox is None:
        eng.compute(left, right)
    else:
        eng.compute(left, right, True, *bbox)
    return eng


@pytest.mark.parametrize("name", list(SMALL_VARIANTS))
def test_small_all_stages_vs_oracle(native, oracle, name):
    prm = variant(configs.params("small"), **SMALL_VARIANTS[name])
    left, right = configs.pair(prm, seed=3)
    ref = oracle.pipeline(prm, left, right)
    eng = run_ours(native, prm, left, right, keep_stages=True)
    assert_stages_equal(eng, prm, ref)
    assert_depth_close(eng.get_ndarray(), ref["out"])
    # the production configuration (no stage materialisation) gives the same final results
    fast = run_ours(native, prm, left, right)
    assert_stages_equal(fast, prm, ref, names=("census0", "census1", "cost", "disp_wta", "disp_right", "disp_med", "depth"))
    assert np.array_equal(fast.get_ndarray().view(np.uint32), eng.get_ndarray().view(np.uint32))


@pytest.mark.parametrize("cfg", ["small435", "C4"])
def test_other_cameras_vs_oracle(native, oracle, cfg):
    prm = configs.params(cfg)
    left, right = configs.pair(prm, seed=11)
    ref = oracle.pipeline(prm, left, right)
    eng = run_ours(native, prm, left, right, keep_stages=True)
    assert_stages_equal(eng, prm, ref)
    assert_depth_close(eng.get_ndarray(), ref["out"])


@pytest.mark.parametrize("geom", [(331, 149, 128), (300, 157, 256), (203, 301, 64), (167, 33, 96)])
def test_odd_geometry_vs_oracle(native, oracle, geom):
    """Sizes that are multiples of nothing: ragged right-most cost blocks (EDGE variant at 16 and 32 columns per
    thread), partial row bands, 2-3 rows per block in the final pass, partial last block."""
    cols, rows, d = geom
    base = configs._sensor_params("D415", max_disp=d, rectified=False, roll_deg=0.5,
                                  scale=(cols, rows, (cols * 3) // 2, (rows * 3) // 2))
    left, right = configs.pair(base, seed=cols)
    ref = oracle.pipeline(base, left, right)
    eng = run_ours(native, base, left, right, keep_stages=True)
    assert_stages_equal(eng, base, ref)
    assert_depth_close(eng.get_ndarray(), ref["out"])
    fast = run_ours(native, base, left, right)
    assert_stages_equal(fast, base, ref, names=("cost", "disp_wta", "disp_right", "disp_med", "depth"))


@pytest.mark.parametrize("cfg,over", [("C1", {}), ("C1", dict(mf_size=7, lr_max_diff=3)), ("C1r", dict(dilation=False)),
                                      ("C3", {}), ("C5", {}), ("odd", {}), ("small", {})])
def test_banded_output_is_bit_identical(native, cfg, over):
    """bind_output(): the final pass runs in column segments and the depth map streams to the bound host buffer
    band by band.  Every stage and the delivered map must equal the unbanded run bit for bit (frames back to back,
    so that stale state of one frame would show in the next)."""
    import torch

    if cfg == "odd":
        prm = configs._sensor_params("D415", max_disp=64, rectified=False, roll_deg=0.5, scale=(523, 211, 700, 300))
    else:
        prm = variant(configs.params(cfg), **over)
    plain = make_engine(native, prm)
    band = make_engine(native, prm)
    out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    band.bind_output(out)
    for seed in (1, 2, 3, 4):  # four frames back to back (they alternate between the engine's lanes)
        left, right = configs.pair(prm, seed=seed)
        plain.compute(left, right)
        out[:] = -7.0
        band.compute(left, right)
        got = band.get_ndarray(out=out)
        assert got is out
        ref = plain.get_ndarray()
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), f"{int((out != ref).sum())} pixels differ"
        for st in ("disp_wta", "disp_right", "disp_med", "depth"):
            a, b = band.get_stage(st), plain.get_stage(st)
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), st
        # other getters still work on a streamed frame
        assert np.array_equal(band.get_ndarray().view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(band.get_cuda().torch().cpu().numpy().view(np.uint32), ref.view(np.uint32))
    band.bind_output(None)
    band.compute(left, right)
    assert np.array_equal(band.get_ndarray().view(np.uint32), ref.view(np.uint32))


def test_banded_output_batched_multi_wave(native):
    """Batched engine whose final pass needs more than one wave of blocks (6 x 256 rows, 8 rows per block): the band
    kernels are gated by ONE waiting warp, so they cannot starve the pass they wait for."""
    import torch

    prm = configs.params("C4")
    n = 6
    pairs = [configs.pair(prm, seed=40 + i) for i in range(n)]
    left = np.stack([p[0] for p in pairs])
    right = np.stack([p[1] for p in pairs])
    plain = make_engine(native, prm, batch=n)
    band = make_engine(native, prm, batch=n)
    out = torch.empty((n, prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    band.bind_output(out)
    for _ in range(3):
        plain.compute(left, right)
        out[:] = -3.0
        band.compute(left, right)
        band.get_ndarray(out=out)
        ref = plain.get_ndarray()
        assert np.array_equal(out.view(np.uint32), ref.reshape(out.shape).view(np.uint32))


@pytest.mark.parametrize("cfg", ["C4", "C1"])
def test_back_to_back_frames_overlap_without_corruption(native, cfg):
    """Frames enqueued back to back on the engine's stream: the front-end of frame k+1 runs on the helper stream while
    frame k is still post-processing (second splat canvas).  Every frame's depth map, snapshotted on the engine stream
    right behind its compute, must equal the map of the same inputs computed alone -- also when host-input, banded and
    ROI frames are mixed into the sequence."""
    import torch

    prm = configs.params(cfg)
    n = 6
    pairs = [configs.pair(prm, seed=70 + i) for i in range(n)]
    solo = make_engine(native, prm)
    want = []
    for l, r in pairs:
        solo.compute(l, r)
        want.append(solo.get_ndarray())
    eng = make_engine(native, prm)
    es = torch.cuda.ExternalStream(eng.cuda_stream)
    dl = [torch.from_numpy(l).cuda() for l, _ in pairs]
    dr = [torch.from_numpy(r).cuda() for _, r in pairs]
    torch.cuda.synchronize()
    snaps = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32, device="cuda") for _ in range(n)]
    for rep in range(2):
        for i in range(n):
            eng.compute(dl[i], dr[i], sync=False)
            with torch.cuda.stream(es):
                snaps[i].copy_(eng.get_cuda().torch(), non_blocking=True)
        es.synchronize()
        for i in range(n):
            assert np.array_equal(snaps[i].cpu().numpy().view(np.uint32), want[i].view(np.uint32)), f"frame {i} (rep {rep})"
    # mixed sequence on one engine: device async, host, host banded, ROI, device async again
    out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
    eng.compute(dl[0], dr[0], sync=False)
    eng.compute(*pairs[1])
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[1].view(np.uint32))
    eng.bind_output(out)
    eng.compute(dl[2], dr[2], sync=False)
    eng.compute(*pairs[3])
    eng.get_ndarray(out=out)
    assert np.array_equal(out.view(np.uint32), want[3].view(np.uint32))
    eng.compute(*pairs[4], True, 16, 8, prm.cols // 2, prm.rows // 2)
    eng.compute(dl[5], dr[5], sync=False)
    eng.compute(*pairs[0])
    eng.get_ndarray(out=out)
    assert np.array_equal(out.view(np.uint32), want[0].view(np.uint32))
    eng.bind_output(None)
    eng.compute(dl[5], dr[5])
    assert np.array_equal(eng.get_ndarray().view(np.uint32), want[5].view(np.uint32))


@pytest.mark.parametrize("bbox", [(8, 4, 64, 40), (0, 0, 96, 64), (31, 23, 33, 37), (60, 30, 36, 34)])
def test_small_bbox_vs_oracle(native, oracle, bbox):
    prm = configs.params("small")
    left, right = configs.pair(prm, seed=5)
    ref = oracle.pipeline(prm, left, right, bbox=bbox)
    eng = run_ours(native, prm, left, right, bbox=bbox, keep_stages=True)
    assert_stages_equal(eng, prm, ref, bbox=bbox)
    assert_depth_close(eng.get_ndarray(), ref["out"])


@pytest.mark.parametrize("case", golden_util.cases())
def test_cuda_engine_vs_reference_golden(native, case):
    """The CUDA engine against the committed outputs of the unmodified reference simsense kernels
    (tests/golden/, captured on a B200): integer stages bit-exact (volumes by SHA-256), float
    stages bit-exact except at the reference's own race pixels, final depth within 1e-4."""
    g = golden_util.Golden(case)
    eng = run_ours(native, g.params, g.left, g.right, bbox=g.bbox, keep_stages=True)
    names = ["census0", "census1", "cost", "L0", "L1", "L2", "LAll", "disp_right", "disp_lr", "disp_med", "depth"]
    stages = {n: get_stage(eng, g.params, n, g.bbox) for n in names}
    golden_util.check_against_golden(g, stages, eng.get_ndarray(), f"cuda[{case}]")
    fast = run_ours(native, g.params, g.left, g.right, bbox=g.bbox)
    assert np.array_equal(fast.get_ndarray().view(np.uint32), eng.get_ndarray().view(np.uint32))


needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libsimsense_ref.so not built")

REF_STAGE_MAP = [("census0", "census0"), ("census1", "census1"), ("cost", "cost"), ("L0", "L0"), ("L1", "L1"),
                 ("L2", "L2"), ("LAll", "LAll"), ("disp_right", "rightDisp")]


# The reference's winnerTakesAll is NOT deterministic for max_disp > 32: its two back-to-back block
# reductions share one static __shared__ scratch array without a barrier in between
# (wta.cu:51,188,192), so on a B200 a few left-disparity pixels per frame come out
# wrong and differ from run to run (measured: 2-9 of 12288 pixels at 128x96/D=64, LAll and the
# right disparity identical).  The integer stages up to LAll and the right disparity must match
# bit-for-bit; for the float disparity and everything downstream of it the reference may deviate
# from the (deterministic, oracle-pinned) result at no more than RACE_FRACTION of the pixels.
RACE_FRACTION = 0.01
# The reference's depthDilation is racy too: it min-writes into neighbours IN PLACE while other
# threads are still reading their own centre value (camera.cu:200-228), so a
# lowered pixel can be propagated a second step.  This engine (and the oracle) implement the
# deterministic snapshot semantics; with dilation on, a few RGB-frame pixels of a reference run may
# therefore differ (measured: 32 of 2 073 600 at C1).


def compare_with_reference(native, prm, left, right, bbox=None):
    ref = RefEngine(prm)
    ref.compute_host(left, right, bbox)
    eng = run_ours(native, prm, left, right, bbox=bbox, keep_stages=True)
    bad = []
    for ours, theirs in REF_STAGE_MAP:
        a = get_stage(eng, prm, ours, bbox)
        b = ref.stage(theirs)
        if not np.array_equal(a, b):
            bad.append(f"{ours}: {int((a != b).sum())} of {a.size} differ")

    def ndiff(ours, theirs):
        a = get_stage(eng, prm, ours, bbox)
        b = ref.stage(theirs)
        return int((a.view(np.uint32) != b.view(np.uint32)).sum()), a.size

    # left disparity: the reference LR-checks in place, so leftDisp is post-LR
    n_lr, size = ndiff("disp_lr", "leftDisp")
    wta_race_possible = prm.max_disp > 32
    if n_lr > (RACE_FRACTION * size if wta_race_possible else 0):
        bad.append(f"disp_lr: {n_lr} of {size} differ (> race allowance)")
    k2 = prm.mf_size ** 2
    if prm.mf_size != 1:
        n, _ = ndiff("disp_med", "filteredDisp")
        if n > k2 * n_lr:
            bad.append(f"disp_med: {n} differ, more than {k2} x {n_lr} race pixels can explain")
    n, _ = ndiff("depth", "depth")
    if n > k2 * n_lr:
        bad.append(f"depth: {n} differ, more than {k2} x {n_lr} race pixels can explain")
    assert not bad, "\n".join(bad)
    ours_depth, ref_depth = eng.get_ndarray(), ref.depth()
    mism = (ours_depth == 0) != (ref_depth == 0)
    both = (ours_depth != 0) & (ref_depth != 0)
    rel = np.zeros(ours_depth.shape)
    rel[both] = np.abs(ours_depth[both] - ref_depth[both]) / ref_depth[both]
    off = int(mism.sum() + (rel > 1e-4).sum())
    print(f"[vs reference] race pixels in leftDisp: {n_lr}; final depth pixels off: {off} of {mism.size}")
    if n_lr == 0 and not prm.dilation:  # no racy stage involved: everything must agree
        assert off == 0
    else:
        allowed = 4 * k2 * n_lr + (0.01 * mism.size if prm.dilation else 0)
        assert off <= allowed, f"{off} final-depth pixels differ from the reference run (allowed {allowed:.0f})"
    ref.close()
    return eng


@needs_ref
@pytest.mark.parametrize("name", ["default", "rectified", "no_lr", "mf1", "mf5", "bf1", "bf3", "census5", "uniq50",
                                  "no_dilation", "d64", "d96", "d128_wider_than_image", "d256", "p_large"])
def test_small_vs_reference_cuda(native, name):
    prm = variant(configs.params("small"), **SMALL_VARIANTS[name])
    left, right = configs.pair(prm, seed=3)
    compare_with_reference(native, prm, left, right)


@needs_ref
def test_small_bbox_vs_reference_cuda(native):
    prm = configs.params("small")
    left, right = configs.pair(prm, seed=5)
    compare_with_reference(native, prm, left, right, bbox=(8, 4, 64, 40))


@needs_ref
def test_c1_full_size_vs_reference_and_oracle(native, oracle):
    """BASELINE configs[0]: single D415-like pair 1280x720, 128 disparities, 4-path SGM + LR +
    median + registration to 1920x1080 -- scalar oracle vs reference simsense vs this engine."""
    prm = configs.params("C1")
    left, right = configs.pair(prm, seed=0)
    eng = compare_with_reference(native, prm, left, right)
    ref = oracle.pipeline(prm, left, right, volumes=False)
    assert_stages_equal(eng, prm, ref, names=("im0", "im1", "census0", "census1", "disp_wta", "disp_right", "disp_lr", "disp_med", "depth"))
    assert_depth_close(eng.get_ndarray(), ref["out"])


@needs_ref
def test_c2_bbox_rgb_point_cloud(native, oracle):
    """BASELINE configs[1]: same pipeline with bbox ROI compute and RGB point-cloud output."""
    import torch

    prm = configs.params("C2")
    left, right = configs.pair(prm, seed=0)
    eng = compare_with_reference(native, prm, left, right, bbox=configs.BBOX_C2)
    rgba = synth.make_rgb(prm.rgb_rows, prm.rgb_cols, 0)
    rgba_t = torch.from_numpy(rgba).cuda()
    pc = eng.get_rgb_point_cloud_ndarray(rgba_t)
    want = oracle.pointcloud(eng.get_ndarray(), rgba, prm.main_fx, prm.main_fy, prm.main_skew, prm.main_cx, prm.main_cy)
    assert pc.shape == (prm.rgb_rows * prm.rgb_cols, 6)
    np.testing.assert_allclose(pc, want, rtol=1e-4, atol=1e-6)
    # the reference's own point-cloud kernel agrees with the oracle formula on the reference's own
    # depth map (its depth may differ from ours at a few race pixels, see RACE_FRACTION)
    ref = RefEngine(prm)
    ref.compute_host(left, right, configs.BBOX_C2)
    ref_pc = ref.rgb_point_cloud(rgba_t.data_ptr())
    want_ref = oracle.pointcloud(ref.depth(), rgba, prm.main_fx, prm.main_fy, prm.main_skew, prm.main_cx, prm.main_cy)
    np.testing.assert_allclose(ref_pc, want_ref, rtol=1e-4, atol=1e-6)
    ref.close()


def test_c3_batched_envs_vs_oracle(native, oracle):
    """BASELINE configs[2] shape (848x480, D=96) with a small batch: each env of one batched call
    equals the oracle run on that env alone."""
    prm = configs.params("C3")
    n = 3
    pairs = [configs.pair(prm, seed=s) for s in range(n)]
    eng = make_engine(native, prm, batch=n)
    eng.compute(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]))
    out = eng.get_ndarray()
    assert out.shape == (n, prm.rgb_rows, prm.rgb_cols)
    for i, (l, r) in enumerate(pairs):
        ref = oracle.pipeline(prm, l, r, volumes=False)
        assert_stages_equal(eng, prm, ref, names=("census0", "disp_wta", "disp_right", "disp_med", "depth"), index=i)
        assert_depth_close(out[i], ref["out"], what=f"env {i}")


def test_c4_many_small_envs_batch_equals_single(native, oracle):
    """BASELINE configs[3] shape (256x256, D=64): a 16-env batch equals 16 single-env computes."""
    prm = configs.params("C4")
    n = 16
    pairs = [configs.pair(prm, seed=100 + s) for s in range(n)]
    eng = make_engine(native, prm, batch=n)
    eng.compute(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]))
    out = eng.get_ndarray()
    single = make_engine(native, prm)
    for i, (l, r) in enumerate(pairs):
        single.compute(l, r)
        assert np.array_equal(single.get_ndarray().view(np.uint32), out[i].view(np.uint32)), f"env {i}"
    ref = oracle.pipeline(prm, *pairs[5], volumes=False)
    assert_depth_close(out[5], ref["out"])


def test_c5_highres_d256_properties(native, oracle):
    """BASELINE configs[4] shape (1920x1080, D=256) at full size: parity with the oracle on the
    final maps plus size-independent properties (idempotence, disparity bounds)."""
    prm = configs.params("C5")
    left, right = configs.pair(prm, seed=0)
    eng = run_ours(native, prm, left, right)
    d = get_stage(eng, prm, "disp_med")
    assert d.max() < prm.max_disp and d.min() >= -1
    first = eng.get_ndarray().copy()
    eng.compute(left, right)
    assert np.array_equal(first.view(np.uint32), eng.get_ndarray().view(np.uint32))
    ref = oracle.pipeline(prm, left, right, volumes=False)
    assert_stages_equal(eng, prm, ref, names=("census0", "census1", "disp_wta", "disp_right", "disp_med", "depth"))
    assert_depth_close(first, ref["out"])


def test_device_rgba_input_equals_host_u8(native, oracle):
    import torch

    prm = configs.params("small")
    left, right = configs.pair(prm, seed=7)
    rl, rr = synth.to_rgba(left), synth.to_rgba(right)
    assert np.array_equal(oracle.float2uint8(rl), left)  # the synthetic RGBA encodes the u8 image exactly
    tl, tr = torch.from_numpy(rl).cuda(), torch.from_numpy(rr).cuda()
    eng = make_engine(native, prm)
    eng.compute(tl, tr)
    a = eng.get_ndarray().copy()
    eng.compute(left, right)
    assert np.array_equal(a.view(np.uint32), eng.get_ndarray().view(np.uint32))
    # out-of-range / negative floats clamp like core.cu:51-58
    weird = rl.copy()
    weird[..., 0].flat[::7] = 1.7
    weird[..., 0].flat[3::11] = -0.3
    rect = make_engine(native, variant(prm, rectified=True))
    rect.compute(torch.from_numpy(weird).cuda(), tr)
    assert np.array_equal(get_stage(rect, prm, "im0"), oracle.float2uint8(weird))
    # device uint8 input (extension)
    eng.compute(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda())
    assert np.array_equal(a.view(np.uint32), eng.get_ndarray().view(np.uint32))


def test_cuda_array_handoff(native):
    """CudaArray contract (unittest/test_sapien/test_cuda_array.py:5-92) + aliasing getters."""
    import torch

    t = torch.tensor([[0, 1, 2], [2, 3, 4]]).float().cuda()
    arr = native.CudaArray(t)
    iface = arr.__cuda_array_interface__
    assert arr.typestr == iface["typestr"] == t.__cuda_array_interface__["typestr"]
    assert tuple(arr.shape) == iface["shape"] == t.__cuda_array_interface__["shape"]
    assert tuple(arr.strides) == iface["strides"] == (12, 4)
    assert arr.ptr == iface["data"][0] == t.data_ptr()
    assert arr.torch().data_ptr() == t.data_ptr()
    t3 = torch.tensor([[0, 1, 2], [2, 3, 4], [3, 4, 5]]).float().cuda()
    sl = t3[1:, :-1]
    arr2 = native.CudaArray(sl)
    assert tuple(arr2.strides) == sl.__cuda_array_interface__["strides"]
    assert arr2.ptr == sl.data_ptr()

    prm = configs.params("small")
    left, right = configs.pair(prm, seed=1)
    eng = make_engine(native, prm)
    with pytest.raises(RuntimeError, match="No computed data stored"):
        eng.get_cuda()
    eng.compute(left, right)
    out = eng.get_cuda()
    assert tuple(out.shape) == (prm.rgb_rows, prm.rgb_cols) and tuple(out.strides) == (4 * prm.rgb_cols, 4)
    assert out.typestr == "f4" and out.cuda_id == torch.cuda.current_device()
    host = eng.get_ndarray()
    via_torch = out.torch()
    via_dlpack = torch.from_dlpack(out.dlpack())
    via_protocol = torch.from_dlpack(out)
    for v in (via_torch, via_dlpack, via_protocol):
        assert v.data_ptr() == out.ptr  # aliases engine memory, no copy
        assert np.array_equal(v.cpu().numpy().view(np.uint32), host.view(np.uint32))
    pc = eng.get_point_cloud_cuda()
    assert tuple(pc.shape) == (prm.rgb_rows * prm.rgb_cols, 3) and tuple(pc.strides) == (12, 4)
    np.testing.assert_array_equal(pc.torch().cpu().numpy(), eng.get_point_cloud_ndarray())


def test_point_cloud_vs_oracle(native, oracle):
    prm = configs.params("small435")
    left, right = configs.pair(prm, seed=2)
    eng = run_ours(native, prm, left, right)
    depth = eng.get_ndarray()
    want = oracle.pointcloud(depth, None, prm.main_fx, prm.main_fy, prm.main_skew, prm.main_cx, prm.main_cy)
    np.testing.assert_allclose(eng.get_point_cloud_ndarray(), want, rtol=1e-4, atol=1e-6)


def test_setters_take_effect_next_frame(native, oracle):
    prm = configs.params("small")
    left, right = configs.pair(prm, seed=9)
    eng = make_engine(native, prm, keep_stages=True)
    eng.compute(left, right)
    eng.set_penalties(4, 60)
    eng.set_census_window_size(5, 9)
    eng.set_matching_block_size(3, 3)
    eng.set_uniqueness_ratio(30)
    eng.set_lr_max_diff(2)
    eng.compute(left, right)
    new = variant(prm, p1=4, p2=60, census_width=5, census_height=9, bf_width=3, bf_height=3, uniq_ratio=30, lr_max_diff=2)
    ref = oracle.pipeline(new, left, right)
    assert_stages_equal(eng, new, ref)
    with pytest.raises(TypeError):
        eng.set_penalties(50, 10)
    with pytest.raises(TypeError):
        eng.set_census_window_size(4, 4)


def test_errors(native):
    prm = configs.params("small")
    left, right = configs.pair(prm, seed=1)
    eng = make_engine(native, prm)
    with pytest.raises(RuntimeError, match="No computed data stored"):
        eng.get_ndarray()
    with pytest.raises(RuntimeError, match="Both images must have the same size"):
        eng.compute(left, right[:-1])
    with pytest.raises(RuntimeError, match="Input image size different from initiated"):
        eng.compute(left[:-1], right[:-1])
    with pytest.raises(TypeError):
        eng.compute(left, right, True, 90, 0, 32, 32)  # bbox leaves the image
    with pytest.raises(TypeError):
        make_engine(native, variant(prm, max_disp=16))
    with pytest.raises(TypeError):
        make_engine(native, variant(prm, mf_size=4))


def test_ir_noise_statistics(native):
    """IR noise (camera.cu:21-75) is statistical: mean ~ I * shape*scale, deterministic per seed,
    different per frame."""
    prm = variant(configs.params("small"), rectified=True, speckle_shape=1333.33, speckle_scale=1 / 1333.33,
                  gaussian_mu=0.0, gaussian_sigma=0.25, mf_size=1)
    img = np.full((prm.rows, prm.cols), 120, np.uint8)
    eng = make_engine(native, prm)
    eng.compute(img, img)
    a = get_stage(eng, prm, "im0").astype(np.float64)
    b = get_stage(eng, prm, "im1").astype(np.float64)
    eng.compute(img, img)
    a2 = get_stage(eng, prm, "im0").astype(np.float64)
    assert abs(a.mean() - 120) < 0.5 and abs(b.mean() - 120) < 0.5
    want_std = np.sqrt((120 ** 2) / 1333.33 + 0.25 ** 2 + 1 / 12)
    assert abs(a.std() - want_std) < 0.35, (a.std(), want_std)
    assert not np.array_equal(a, b) and not np.array_equal(a, a2)
    eng2 = make_engine(native, prm)
    eng2.compute(img, img)
    assert np.array_equal(get_stage(eng2, prm, "im0"), a.astype(np.uint8))
