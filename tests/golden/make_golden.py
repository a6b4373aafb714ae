"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference simsense CUDA code.

Run on a GPU box (the reference has no CPU path), after `make -C oracle ref` was done where
/root/reference exists (the built oracle/_ref/libsimsense_ref.so travels with the snapshot):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
    cp gpurun_out/golden/* tests/golden/

Each case file <case>.npz holds everything a checker needs WITHOUT numpy-RNG / OpenCV / the
reference being present: the engine parameters (JSON), the calibration planes, the input pair, and
the reference's own per-stage outputs read through the harness subclass (oracle/ref_harness.cu;
protected members of simsense/core.h:83-93):

    2-D stages (census0/1 u32, leftDisp f32 post-LR, rightDisp u16, filteredDisp f32, depth f32,
                rgbDepth f32)                                   -> stored as arrays
    volumes    (rawcost, hsum, cost, L0, L1, L2, LAll; u16)     -> stored as SHA-256 of the bytes
                                                                   (+ LAll itself for the tiny case)

The reference's winnerTakesAll races for max_disp > 32 (two block reductions share one static
__shared__ scratch array, wta.cu:51,188,192) and its in-place depthDilation races too
(camera.cu:200-228), so float stages are captured RUNS times and the per-pixel majority is stored
together with the number of pixels that were not unanimous (`*_unstable`).
"""
from __future__ import annotations

import dataclasses
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Params, RefEngine, configs  # noqa: E402

RUNS = 7
VOLUMES = ("rawcost", "hsum", "cost", "L0", "L1", "L2", "LAll")
PLANAR = ("census0", "census1", "rightDisp")
RACY = ("leftDisp", "filteredDisp", "depth", "rgbDepth")

# case name -> (config, overrides, pair seed, bbox)
CASES = {
    "small_default": ("small", {}, 3, None),
    "small_rectified_nodil": ("small", dict(rectified=True, dilation=False), 3, None),
    "small_bbox": ("small", {}, 5, (8, 4, 64, 40)),
    "small_bf1_mf5_lr0": ("small", dict(bf_width=1, bf_height=1, mf_size=5, lr_max_diff=0), 4, None),
    "small_bf3_census5_uniq50": ("small", dict(bf_width=3, bf_height=3, census_width=5, census_height=5, uniq_ratio=50), 6, None),
    "small_census9x7_nolr_mf1": ("small", dict(census_width=9, census_height=7, lr_max_diff=255, mf_size=1), 7, None),
    "small_d96_plarge": ("small", dict(max_disp=96, p1=100, p2=223), 8, None),
    "small_d33": ("small", dict(max_disp=33), 9, None),
    "small435_d64": ("small435", {}, 11, None),
    "c4_256x256_d64": ("C4", {}, 1, None),
}
SCALARS = [f.name for f in dataclasses.fields(Params) if f.name not in ("map_lx", "map_ly", "map_rx", "map_ry", "a1", "a2", "a3")]


def majority(stack: np.ndarray):
    """Per-pixel most frequent bit pattern of a [runs, ...] float stack and the unstable count."""
    bits = stack.view(np.uint32)
    srt = np.sort(bits, axis=0)
    med = srt[len(srt) // 2]  # with > half of the runs agreeing the median IS the majority
    unstable = int((bits != bits[0]).any(axis=0).sum())
    return med.view(np.float32), unstable


def main(out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    manifest = {}
    for case, (cfg, over, seed, bbox) in CASES.items():
        prm = dataclasses.replace(configs.params(cfg), **over)
        left, right = configs.pair(prm, seed=seed)
        planes = prm.planes()
        data = {"left": left, "right": right, **{k: v.reshape(prm.rows, prm.cols) for k, v in planes.items()}}
        meta = {"params": {k: (getattr(prm, k) if not isinstance(getattr(prm, k), (np.floating, np.integer)) else getattr(prm, k).item()) for k in SCALARS},
                "bbox": bbox, "seed": seed, "config": cfg, "runs": RUNS, "sha256": {}, "unstable": {}}
        racy = {k: [] for k in RACY}
        for run in range(RUNS):
            ref = RefEngine(prm)
            ref.compute_host(left, right, bbox)
            if run == 0:
                for name in VOLUMES:
                    try:
                        v = ref.stage(name)
                    except RuntimeError:
                        continue
                    meta["sha256"][name] = hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()
                    if name == "LAll" and v.nbytes <= 512 * 1024:
                        data["LAll"] = v.copy()
                for name in PLANAR:
                    data[name] = ref.stage(name).copy()
            else:
                for name in PLANAR:
                    assert np.array_equal(data[name], ref.stage(name)), f"{case}: integer stage {name} not repeatable"
            for name in RACY:
                if name == "filteredDisp" and prm.mf_size == 1:
                    continue
                racy[name].append((ref.depth() if name == "rgbDepth" else ref.stage(name)).copy())
            ref.close()
        for name, runs in racy.items():
            if runs:
                data[name], meta["unstable"][name] = majority(np.stack(runs))
        np.savez_compressed(os.path.join(out_dir, case + ".npz"), meta=np.frombuffer(json.dumps(meta).encode(), np.uint8), **data)
        manifest[case] = {"unstable": meta["unstable"], "bytes": os.path.getsize(os.path.join(out_dir, case + ".npz"))}
        print(case, manifest[case], flush=True)
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "unmodified reference simsense (3rd_party/simsense @ 9340b6069ce7) recompiled for sm_100a, run on a B200",
                   "cases": manifest}, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
