"""Seeded random configurations: every stage of the CUDA engine against the oracle (bit-exact integer stages,
<= 1e-4 relative depth).  Sizes, disparity ranges, window sizes, penalties and switches are drawn together so that
combinations nobody enumerated by hand get exercised (fast packed-u16 path, generic fallback, ragged tiles, ROI)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import configs
from tests.common import assert_depth_close, assert_stages_equal, make_engine

pytestmark = pytest.mark.gpu


def _draw(seed: int):
    rng = np.random.default_rng(1000 + seed)
    cols = int(rng.integers(48, 340))
    rows = int(rng.integers(40, 200))
    d = int(rng.choice([32, 40, 48, 64, 72, 96, 104, 128, 136, 160, 192, 256, 33, 57, 130]))
    cw = int(rng.choice([3, 5, 7, 7, 7, 9]))
    ch = int(rng.choice([3, 5, 7, 7, 7]))
    bw = int(rng.choice([1, 3, 5, 7, 7, 7, 9]))
    bh = bw if rng.random() < 0.8 else int(rng.choice([1, 3, 5]))
    p1 = int(rng.integers(1, 60))
    p2 = int(rng.integers(p1 + 1, 224))
    over = dict(census_width=cw, census_height=ch, bf_width=bw, bf_height=bh, p1=p1, p2=p2,
                uniq_ratio=int(rng.choice([0, 5, 15, 15, 40, 100, 200])), lr_max_diff=int(rng.choice([0, 1, 1, 2, 255])),
                mf_size=int(rng.choice([1, 3, 3, 5, 7])), dilation=bool(rng.random() < 0.6))
    rectified = bool(rng.random() < 0.4)
    prm = configs._sensor_params("D415" if rng.random() < 0.5 else "D435", max_disp=d, rectified=rectified,
                                 roll_deg=0.0 if rectified else 0.5,
                                 scale=(cols, rows, int(cols * rng.choice([1.0, 1.5])), int(rows * rng.choice([1.0, 1.5]))), **over)
    bbox = None
    if rng.random() < 0.25:
        w = int(rng.integers(max(bw, 33), cols + 1))
        h = int(rng.integers(max(bh, 33), rows + 1))
        bbox = (int(rng.integers(0, cols - w + 1)), int(rng.integers(0, rows - h + 1)), w, h)
    return prm, bbox


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_vs_oracle(native, oracle, seed):
    prm, bbox = _draw(seed)
    left, right = configs.pair(prm, seed=seed)
    ref = oracle.pipeline(prm, left, right, bbox=bbox)
    eng = make_engine(native, prm, keep_stages=True)
    if bbox is None:
        eng.compute(left, right)
    else:
        eng.compute(left, right, True, *bbox)
    assert_stages_equal(eng, prm, ref, bbox=bbox)
    assert_depth_close(eng.get_ndarray(), ref["out"])
    fast = make_engine(native, prm)
    if bbox is None:
        fast.compute(left, right)
    else:
        fast.compute(left, right, True, *bbox)
    assert_stages_equal(fast, prm, ref, bbox=bbox, names=("cost", "disp_wta", "disp_right", "disp_med", "depth"))
    assert np.array_equal(fast.get_ndarray().view(np.uint32), eng.get_ndarray().view(np.uint32))
