"""Helpers shared by the parity tests."""
from __future__ import annotations

import dataclasses

import numpy as np

from oracle import Params, configs

# stage name in our engine -> (oracle key, reference-harness key)
STAGES = [
    ("im0", "im0", None), ("im1", "im1", None),
    ("census0", "census0", "census0"), ("census1", "census1", "census1"),
    ("cost", "cost", "cost"), ("L0", "L0", "L0"), ("L1", "L1", "L1"), ("L2", "L2", "L2"),
    ("L3", "L3", None), ("LAll", "LAll", "LAll"),
    ("disp_wta", "disp_wta", None), ("disp_right", "disp_right", "rightDisp"),
    ("disp_lr", "disp_lr", "leftDisp"), ("disp_med", "disp_med", None),
    ("disp_full", "disp_full", None), ("depth", "depth", "depth"),
]


def make_engine(native, prm: Params, **kw):
    return native.DepthSensorEngine(*prm.engine_args(), **kw)


def variant(prm: Params, **over) -> Params:
    return dataclasses.replace(prm, **over)


def stage_shape(prm: Params, name: str, bbox=None):
    rows, cols = (prm.rows, prm.cols) if bbox is None else (bbox[3], bbox[2])
    if name in ("cost", "L0", "L1", "L2", "L3", "LAll"):
        return (rows, cols, prm.max_disp)
    if name in ("disp_full", "depth"):
        return (prm.rows, prm.cols)
    return (rows, cols)


def get_stage(eng, prm, name, bbox=None, index=0):
    return eng.get_stage(name, index).reshape(stage_shape(prm, name, bbox))


def assert_stages_equal(eng, prm, ref: dict, bbox=None, names=None, index=0):
    """Bit-exact comparison of every stage of `eng` against the dict `ref` (oracle keys)."""
    bad = []
    for ours, okey, _ in STAGES:
        if names is not None and ours not in names:
            continue
        if okey not in ref:
            continue
        a = get_stage(eng, prm, ours, bbox, index)
        b = ref[okey]
        if a.dtype.kind == "f":
            same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
        else:
            same = np.array_equal(a, b)
        if not same:
            n = int((a != b).sum())
            idx = np.argwhere(a != b)[:3].tolist()
            bad.append(f"{ours}: {n} of {a.size} differ, first at {idx}: ours={[a[tuple(i)] for i in idx]} ref={[b[tuple(i)] for i in idx]}")
    assert not bad, "stage mismatches:\n" + "\n".join(bad)


def assert_depth_close(a: np.ndarray, b: np.ndarray, rtol=1e-4, what="depth"):
    """<=1e-4 relative on depth (north_star tolerance); zero/non-zero pattern must agree."""
    a = np.asarray(a, np.float32).reshape(b.shape)
    za, zb = a == 0, b == 0
    assert np.array_equal(za, zb), f"{what}: validity masks differ at {int((za != zb).sum())} pixels"
    nz = ~zb
    if nz.any():
        rel = np.abs(a[nz] - b[nz]) / np.abs(b[nz])
        assert rel.max() <= rtol, f"{what}: max rel err {rel.max():.3e} > {rtol}"
