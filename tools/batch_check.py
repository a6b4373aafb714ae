#!/usr/bin/env python
"""Sanitizer target for the batched (throughput-regime) kernels: python tools/batch_check.py small435 16
One batch through the production engine, every integer stage the debug engine keeps compared bit by bit with
the production engine's outputs (the debug engine runs the one-row final pass, the production one the half-warp
kernel at D = 64 / 96), twice, plus a second batch with other images in between."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import configs
from sapien_b200 import simsense

cfg, batch = sys.argv[1], int(sys.argv[2])
prm = configs.params(cfg)
pairs = [configs.pair(prm, seed=300 + i) for i in range(batch + 1)]
la, ra = np.stack([p[0] for p in pairs[:batch]]), np.stack([p[1] for p in pairs[:batch]])
lb, rb = np.stack([p[0] for p in pairs[1:]]), np.stack([p[1] for p in pairs[1:]])
dbg = simsense.DepthSensorEngine(*prm.engine_args(), batch=batch, keep_stages=True)
eng = simsense.DepthSensorEngine(*prm.engine_args(), batch=batch)
bad = 0
for l, r in ((la, ra), (lb, rb), (la, ra)):
    dbg.compute(l, r)
    eng.compute(l, r)
    a, b = dbg.get_ndarray(), eng.get_ndarray()
    bad += int(not np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    for st in ("disp_right", "disp_med", "depth"):
        x, y = dbg.get_stage(st), eng.get_stage(st)
        bad += int(not np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)))
print(f"{cfg} x {batch}: 3 batches, {bad} mismatches")
sys.exit(1 if bad else 0)
