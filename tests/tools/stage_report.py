"""Debug aid: per-stage mismatch counts of the CUDA engine against the oracle (and the reference
CUDA build when present) for a few configurations.  Run on the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import Oracle, REF_SO, RefEngine, configs  # noqa: E402
from sapien_b200 import simsense  # noqa: E402
from tests.common import STAGES, get_stage, make_engine, variant  # noqa: E402

names = sys.argv[1:] or ["small"]
orc = Oracle()
for cfg in names:
    over = {}
    if ":" in cfg:
        cfg, kv = cfg.split(":", 1)
        over = {k: int(v) for k, v in (x.split("=") for x in kv.split(","))}
    prm = variant(configs.params(cfg), **over)
    left, right = configs.pair(prm, seed=3)
    big = prm.rows * prm.cols > 300000
    ref = orc.pipeline(prm, left, right, volumes=not big)
    eng = make_engine(simsense, prm, keep_stages=True)
    eng.compute(left, right)
    print(f"== {cfg} {over} {prm.rows}x{prm.cols} D={prm.max_disp}")
    for ours, okey, rkey in STAGES:
        if okey not in ref:
            continue
        a = get_stage(eng, prm, ours)
        b = ref[okey]
        neq = a != b
        msg = f"  {ours:10s} vs oracle: {int(neq.sum()):9d} / {a.size} differ"
        if neq.any():
            i = tuple(np.argwhere(neq)[0])
            msg += f"  first {i}: ours={a[i]} oracle={b[i]}"
        print(msg)
    out = eng.get_ndarray()
    print("  final depth: validity mismatches", int(((out == 0) != (ref['out'] == 0)).sum()),
          "max rel", float(np.nanmax(np.abs(out - ref['out']) / np.maximum(ref['out'], 1e-9))))
    if os.path.exists(REF_SO):
        r = RefEngine(prm)
        r.compute_host(left, right)
        for ours, okey, rkey in STAGES:
            if rkey is None or (big and okey not in ref and ours not in ("cost", "L0", "L1", "L2", "LAll")):
                continue
            b = r.stage(rkey)
            o = ref.get(okey)
            if o is not None:
                print(f"  {rkey:10s} reference vs oracle: {int((b != o).sum()):9d} differ")
            else:
                a = get_stage(eng, prm, ours)
                print(f"  {rkey:10s} reference vs ours  : {int((b != a).sum()):9d} differ")
        rd = r.depth()
        print("  rgbDepth reference vs oracle: validity mismatches", int(((rd == 0) != (ref['out'] == 0)).sum()),
              "max rel", float(np.nanmax(np.abs(rd - ref['out']) / np.maximum(ref['out'], 1e-9))))
        r.close()
