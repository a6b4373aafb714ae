"""Pins the C oracle (oracle/simsense_oracle.c) WITHOUT a GPU:

1. against the golden fixtures in tests/golden/ -- per-stage outputs of the UNMODIFIED reference
   simsense CUDA kernels captured on a B200 by tests/golden/make_golden.py;
2. against an independent numpy restatement of SURVEY.md Appendix A written in this file
   (different language, different loop structure) on small inputs;
3. against hand-computable known answers.
"""
import numpy as np
import pytest

from oracle import Params, configs
from tests import golden_util
from tests.common import variant

CASES = golden_util.cases()


@pytest.mark.skipif(not CASES, reason="no golden fixtures generated yet")
@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(oracle, case):
    g = golden_util.Golden(case)
    got = oracle.pipeline(g.params, g.left, g.right, bbox=g.bbox)
    off = golden_util.check_against_golden(g, got, got["out"], f"oracle[{case}]")
    print(f"{case}: final-depth pixels off vs reference run: {off}; reference-unstable pixels: {g.unstable}")


def test_golden_fixtures_exist():
    assert len(CASES) >= 8, "tests/golden/*.npz missing: run tests/golden/make_golden.py on a GPU box"


# ------------------------------------------------------------------ independent numpy restatement
def np_census(img, cw, ch):
    h, w = img.shape
    left, top = (cw - 1) // 2, (ch - 1) // 2
    pad = np.zeros((h + 2 * top, w + 2 * left), np.int32)
    pad[top:top + h, left:left + w] = img
    out = np.zeros((h, w), np.uint64)
    for i in range(top + 1):
        for j in range(cw // 2 if i == top else cw):
            a = pad[i:i + h, j:j + w]
            b = pad[2 * top - i:2 * top - i + h, 2 * left - j:2 * left - j + w]
            out |= (a >= b).astype(np.uint64) << np.uint64(i * cw + j)
    return (out & np.uint64(0xffffffff)).astype(np.uint32)


def np_cost(cl, cr, D, bw, bh):
    h, w = cl.shape
    xs = np.arange(w)
    ham = np.empty((h, w, D), np.int64)
    for d in range(D):
        x = cl ^ cr[:, np.maximum(xs - d, 0)]
        ham[:, :, d] = np.array([bin(v).count("1") for v in x.ravel()]).reshape(h, w)
    raw = ham.copy()
    if bw * bh != 1:
        p = np.pad(ham, ((0, 0), (bw // 2, bw // 2), (0, 0)), mode="edge")
        ham = sum(p[:, k:k + w] for k in range(bw))
        p = np.pad(ham, ((bh // 2, bh // 2), (0, 0), (0, 0)), mode="edge")
        ham = sum(p[k:k + h] for k in range(bh))
    return raw.astype(np.uint16), ham.astype(np.uint16)


def np_path(C, P1, P2, dy, dx):
    h, w, D = C.shape
    C = C.astype(np.int64)
    L = np.zeros_like(C)
    ys = range(h) if dy >= 0 else range(h - 1, -1, -1)
    xs = range(w) if dx >= 0 else range(w - 1, -1, -1)
    big = 1 << 40
    for y in ys:
        for x in xs:
            py, px = y - dy, x - dx
            if not (0 <= py < h and 0 <= px < w):
                L[y, x] = C[y, x]
                continue
            prev = L[py, px]
            m = prev.min()
            lo = np.concatenate(([big], prev[:-1])) + P1
            hi = np.concatenate((prev[1:], [big])) + P1
            L[y, x] = C[y, x] + np.minimum(np.minimum(prev, lo), np.minimum(hi, m + P2)) - m
    return L


def np_wta(LA, uniq):
    h, w, D = LA.shape
    LA = LA.astype(np.int64)
    dl = np.empty((h, w), np.float32)
    dr = np.empty((h, w), np.uint16)
    ds = np.arange(D)
    for y in range(h):
        for x in range(w):
            v = LA[y, x]
            d = int(v.argmin())
            m = int(v[d])
            ok = np.all((v * (100 - uniq) >= m * 100) | (np.abs(ds - d) <= 1))
            if not ok:
                dl[y, x] = -1.0
            elif 0 < d < D - 1:
                y0, y2 = int(v[d - 1]), int(v[d + 1])
                dl[y, x] = np.float32(d) - np.float32(np.float64(y2 - y0) / (2.0 * np.float64(y0 - 2 * m + y2)))
            else:
                dl[y, x] = d
            n = min(D, w - x)
            diag = LA[y, x + ds[:n], ds[:n]]
            dr[y, x] = int(diag.argmin())
    return dl, dr


@pytest.mark.parametrize("over", [dict(), dict(bf_width=3, bf_height=1, census_width=5, census_height=3, uniq_ratio=40, p1=3, p2=50),
                                  dict(bf_width=1, bf_height=1, census_width=9, census_height=7, uniq_ratio=0)])
def test_oracle_matches_independent_numpy_restatement(oracle, over):
    prm = variant(Params(rows=36, cols=44, rgb_rows=36, rgb_cols=44, focal_len=100.0, baseline_len=0.05,
                         max_disp=32, rectified=True, lr_max_diff=255, mf_size=1, dilation=False), **over)
    from sapien_b200.synth import make_pair

    left, right, _ = make_pair(prm.rows, prm.cols, 24, seed=21)
    got = oracle.pipeline(prm, left, right)
    cl, cr = np_census(left, prm.census_width, prm.census_height), np_census(right, prm.census_width, prm.census_height)
    assert np.array_equal(got["census0"], cl) and np.array_equal(got["census1"], cr)
    raw, cost = np_cost(cl, cr, prm.max_disp, prm.bf_width, prm.bf_height)
    if prm.bf_width * prm.bf_height != 1:
        assert np.array_equal(got["rawcost"], raw)
    assert np.array_equal(got["cost"], cost)
    area = prm.bf_width * prm.bf_height
    P1, P2 = prm.p1 * area, prm.p2 * area
    Ls = [np_path(cost, P1, P2, 0, 1), np_path(cost, P1, P2, 0, -1), np_path(cost, P1, P2, 1, 0), np_path(cost, P1, P2, -1, 0)]
    for name, L in zip(("L0", "L1", "L2", "L3"), Ls):
        assert np.array_equal(got[name], L.astype(np.uint16)), name
    la = (sum(Ls) // 4).astype(np.uint16)
    assert np.array_equal(got["LAll"], la)
    dl, dr = np_wta(la, prm.uniq_ratio)
    assert np.array_equal(got["disp_right"], dr)
    assert np.array_equal(got["disp_wta"].view(np.uint32), dl.view(np.uint32))


# ------------------------------------------------------------------------------- known answers
def test_known_answer_constant_shift(oracle):
    """A textured image shifted by exactly 9 px must come out as disparity 9 (and z = f*b/9)
    away from the borders."""
    rng = np.random.default_rng(5)
    prm = Params(rows=48, cols=96, rgb_rows=48, rgb_cols=96, focal_len=200.0, baseline_len=0.05, max_disp=32,
                 rectified=True, dilation=False, min_depth=0.01, max_depth=100.0)
    left = rng.integers(0, 256, (48, 96), dtype=np.uint8)
    right = np.roll(left, -9, axis=1)
    got = oracle.pipeline(prm, left, right, registration=False)
    core = got["disp_med"][8:-8, 48:-16]
    assert np.all(np.abs(core - 9.0) < 0.02)  # integer winner 9, parabola offset ~1e-3 on random texture
    z = got["out"][8:-8, 48:-16]
    assert np.array_equal(z, (np.float32(200.0) * np.float32(0.05)) / core)  # (f*b)/d, camera.cu:167


def test_known_answer_census_bits_of_flat_image(oracle):
    """Flat image: every comparison a >= b is true inside the image; at the image corner the
    padded zeros make (0 >= v) false for window positions that fall outside."""
    prm = Params(rows=32, cols=32, rgb_rows=32, rgb_cols=32, focal_len=1.0, baseline_len=1.0, max_disp=32, rectified=True)
    img = np.full((32, 32), 77, np.uint8)
    got = oracle.pipeline(prm, img, img)
    assert got["census0"][16, 16] == (1 << 24) - 1  # 7x7 -> 24 comparisons
    # top-left pixel: a = P(-3+i, -3+j) is always outside (0); b = P(3-i, 3-j) is outside too for
    # j >= 4 (0 >= 0 -> bit set), inside (77) otherwise -> bits i*7+j, i in 0..2, j in 4..6
    assert got["census0"][0, 0] == sum(1 << (i * 7 + j) for i in range(3) for j in (4, 5, 6))
    # bottom-right pixel: a is inside only when i=3 (j<3: x-3+j inside), b is outside for those -> bits 21..23 set
    # and for i<3: a inside (y-3+i <= y), b = P(y+3-i, .) outside -> set.  All 24 set.
    assert got["census0"][31, 31] == (1 << 24) - 1
    assert np.all(got["cost"] == 0) or got["cost"].max() <= 24 * 49


def test_known_answer_sgm_on_flat_costs_and_depth_rules(oracle):
    prm = Params(rows=32, cols=40, rgb_rows=32, rgb_cols=40, focal_len=50.0, baseline_len=0.1, max_disp=32, rectified=True,
                 lr_max_diff=255, mf_size=1, dilation=False)
    img = np.zeros((32, 40), np.uint8)
    got = oracle.pipeline(prm, img, img, registration=False)
    # all census codes equal -> all costs 0 -> every path cost 0 -> argmin 0 (lowest-d tie-break), unique fails
    assert not got["cost"].any() and not got["LAll"].any()
    assert np.all(got["disp_right"] == 0)
    # uniqueness: LAll(d)*(100-15) >= 0 holds for every d -> unique, d*=0 -> disparity 0 -> depth 0 (d<=0 rule)
    assert np.all(got["disp_wta"] == 0.0) and np.all(got["depth"] == 0.0) and np.all(got["out"] == 0.0)
