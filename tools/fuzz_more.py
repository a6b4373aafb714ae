#!/usr/bin/env python
"""Extended fuzzing beyond tests/test_fuzz_gpu.py (needs a B200): seeds [a, b) of the same generator, each configuration as ONE
frame and as a BATCH in the throughput regime (half-warp final pass at D = 64 / 96, multi-wave cost grids), production kernels
against the oracle on cost (single frame), both disparities, median and depth.    python tools/fuzz_more.py 24 224"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import Oracle, configs
from sapien_b200 import simsense
from tests.common import assert_stages_equal, make_engine
from tests.test_fuzz_gpu import _draw

a, b = int(sys.argv[1]), int(sys.argv[2])
oracle = Oracle()
bad = 0
for seed in range(a, b):
    prm, bbox = _draw(seed)
    try:
        left, right = configs.pair(prm, seed=seed)
        ref = oracle.pipeline(prm, left, right, bbox=bbox)
        fast = make_engine(simsense, prm)
        if bbox is None:
            fast.compute(left, right)
        else:
            fast.compute(left, right, True, *bbox)
        assert_stages_equal(fast, prm, ref, bbox=bbox, names=("cost", "disp_wta", "disp_right", "disp_med", "depth"))
        # batch: enough rows for the throughput regime of the final pass
        rows = prm.rows if bbox is None else bbox[3]
        n = 5 * 148 // rows + 2
        pairs = [(left, right)] + [configs.pair(prm, seed=10_000 + seed * 7 + i) for i in range(1, min(n, 3))]
        ls = np.stack([pairs[i % len(pairs)][0] for i in range(n)])
        rs = np.stack([pairs[i % len(pairs)][1] for i in range(n)])
        eng = make_engine(simsense, prm, batch=n)
        if bbox is None:
            eng.compute(ls, rs)
        else:
            eng.compute(ls, rs, True, *bbox)
        refs = [ref] + [oracle.pipeline(prm, l, r, bbox=bbox, volumes=False) for l, r in pairs[1:]]
        for i in sorted({0, 1, n // 2, n - 1}):
            assert_stages_equal(eng, prm, refs[i % len(pairs)], bbox=bbox, names=("disp_wta", "disp_right", "disp_med", "depth"), index=i)
    except AssertionError as ex:
        bad += 1
        print(f"seed {seed}: MISMATCH {prm.cols}x{prm.rows} D={prm.max_disp} bf={prm.bf_width}x{prm.bf_height} bbox={bbox}: {str(ex)[:300]}", flush=True)
    except Exception as ex:  # configurations the engine refuses are refused by the reference rules as well
        print(f"seed {seed}: {type(ex).__name__}: {str(ex)[:160]}", flush=True)
print(f"seeds {a}..{b - 1}: {bad} mismatches")
sys.exit(1 if bad else 0)
