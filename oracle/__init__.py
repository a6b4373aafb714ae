"""ctypes bindings of the CHECKERS -- test infrastructure only.

* ``Oracle``    : oracle/liboracle.so, the scalar C restatement (oracle/simsense_oracle.c).
* ``RefEngine`` : oracle/_ref/libsimsense_ref.so, the unmodified reference simsense CUDA code
                  behind the shim oracle/ref_harness.cu (needs a GPU).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package.  Nothing under sapien_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsimsense_ref.so")
REF_TIMED_SO = os.path.join(HERE, "_ref", "libsimsense_ref_timed.so")  # same sources, -DPRINT_RUNTIME


def build(ref: bool = True) -> None:
    """Builds the checkers (oracle/Makefile).  The reference part is skipped when
    /root/reference is absent (GPU box: the prebuilt .so travels with the snapshot)."""
    targets = ["liboracle.so"] + (["ref"] if ref else [])
    r = subprocess.run(["make", "-C", HERE, *targets], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


@dataclass
class Params:
    """All engine parameters in one place (reference ctor order, core.h:43-55)."""
    rows: int
    cols: int
    rgb_rows: int
    rgb_cols: int
    focal_len: float
    baseline_len: float
    min_depth: float = 0.2
    max_depth: float = 10.0
    ir_noise_seed: int = 0
    speckle_shape: float = 0.0
    speckle_scale: float = 0.0
    gaussian_mu: float = 0.0
    gaussian_sigma: float = 0.0
    rectified: bool = True
    census_width: int = 7
    census_height: int = 7
    max_disp: int = 128
    bf_width: int = 7
    bf_height: int = 7
    p1: int = 8
    p2: int = 32
    uniq_ratio: int = 15
    lr_max_diff: int = 1
    mf_size: int = 3
    map_lx: Optional[np.ndarray] = None
    map_ly: Optional[np.ndarray] = None
    map_rx: Optional[np.ndarray] = None
    map_ry: Optional[np.ndarray] = None
    a1: Optional[np.ndarray] = None
    a2: Optional[np.ndarray] = None
    a3: Optional[np.ndarray] = None
    b1: float = 0.0
    b2: float = 0.0
    b3: float = 0.0
    dilation: bool = True
    main_fx: float = 1.0
    main_fy: float = 1.0
    main_skew: float = 0.0
    main_cx: float = 0.0
    main_cy: float = 0.0

    def planes(self):
        n = self.rows * self.cols
        def f(a, fill):
            if a is None:
                a = np.full((self.rows, self.cols), fill, dtype=np.float32)
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == n
            return a
        if self.map_lx is None:
            xs, ys = np.meshgrid(np.arange(self.cols, dtype=np.float32), np.arange(self.rows, dtype=np.float32))
            ident = (xs, ys)
        else:
            ident = (None, None)
        return dict(
            map_lx=f(self.map_lx if self.map_lx is not None else ident[0], 0),
            map_ly=f(self.map_ly if self.map_ly is not None else ident[1], 0),
            map_rx=f(self.map_rx if self.map_rx is not None else ident[0], 0),
            map_ry=f(self.map_ry if self.map_ry is not None else ident[1], 0),
            a1=f(self.a1, 0), a2=f(self.a2, 0), a3=f(self.a3, 1))

    def engine_args(self):
        """Positional arguments of DepthSensorEngine (python/pybind/simsense.cpp:54-73)."""
        p = self.planes()
        return (self.rows, self.cols, self.rgb_rows, self.rgb_cols, self.focal_len, self.baseline_len,
                self.min_depth, self.max_depth, self.ir_noise_seed, self.speckle_shape, self.speckle_scale,
                self.gaussian_mu, self.gaussian_sigma, self.rectified, self.census_width, self.census_height,
                self.max_disp, self.bf_width, self.bf_height, self.p1, self.p2, self.uniq_ratio,
                self.lr_max_diff, self.mf_size, p["map_lx"], p["map_ly"], p["map_rx"], p["map_ry"],
                p["a1"], p["a2"], p["a3"], self.b1, self.b2, self.b3, self.dilation, self.main_fx,
                self.main_fy, self.main_skew, self.main_cx, self.main_cy)


class _OrcCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("rows", "cols", "rgb_rows", "rgb_cols")] + \
               [(n, C.c_float) for n in ("focal", "baseline", "min_depth", "max_depth")] + \
               [(n, C.c_int) for n in ("rectified", "census_w", "census_h", "max_disp", "bf_w", "bf_h", "p1", "p2",
                                       "uniq", "lr_max_diff", "mf_size")] + \
               [(n, C.c_float) for n in ("b1", "b2", "b3")] + \
               [(n, C.c_int) for n in ("dilation", "registration", "bbox", "bbox_x", "bbox_y", "bbox_w", "bbox_h")]


_STAGE_FIELDS = [("im0", np.uint8), ("im1", np.uint8), ("census0", np.uint32), ("census1", np.uint32),
                 ("rawcost", np.uint16), ("cost", np.uint16), ("L0", np.uint16), ("L1", np.uint16),
                 ("L2", np.uint16), ("L3", np.uint16), ("LAll", np.uint16), ("disp_wta", np.float32),
                 ("disp_lr", np.float32), ("disp_med", np.float32), ("disp_right", np.uint16),
                 ("disp_full", np.float32), ("depth", np.float32)]


class _OrcStages(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in _STAGE_FIELDS]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Scalar CPU restatement; `pipeline` returns the final depth and (optionally) every stage."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_pipeline.restype = C.c_int
        self.lib.orc_pipeline.argtypes = [C.POINTER(_OrcCfg)] + [C.c_void_p] * 9 + [C.c_void_p, C.POINTER(_OrcStages)]

    def pipeline(self, prm: Params, left: np.ndarray, right: np.ndarray, bbox=None, stages=True,
                 volumes=True, registration=True) -> Dict[str, np.ndarray]:
        left = np.ascontiguousarray(left, dtype=np.uint8)
        right = np.ascontiguousarray(right, dtype=np.uint8)
        assert left.shape == (prm.rows, prm.cols) == right.shape
        cfg = _OrcCfg(prm.rows, prm.cols, prm.rgb_rows, prm.rgb_cols, prm.focal_len, prm.baseline_len,
                      prm.min_depth, prm.max_depth, int(prm.rectified), prm.census_width, prm.census_height,
                      prm.max_disp, prm.bf_width, prm.bf_height, prm.p1, prm.p2, prm.uniq_ratio,
                      255 if prm.lr_max_diff == -1 else prm.lr_max_diff, prm.mf_size, prm.b1, prm.b2, prm.b3,
                      int(prm.dilation), int(registration), 0, 0, 0, 0, 0)
        rows, cols = prm.rows, prm.cols
        if bbox is not None:
            cfg.bbox, cfg.bbox_x, cfg.bbox_y, cfg.bbox_w, cfg.bbox_h = 1, *bbox
            rows, cols = bbox[3], bbox[2]
        pl = prm.planes()
        out_shape = (prm.rgb_rows, prm.rgb_cols) if registration else (prm.rows, prm.cols)
        out = {"out": np.empty(out_shape, np.float32)}
        st = _OrcStages()
        if stages:
            msz, fsz, D = rows * cols, prm.rows * prm.cols, prm.max_disp
            for name, dt in _STAGE_FIELDS:
                is_vol = name in ("rawcost", "cost", "L0", "L1", "L2", "L3", "LAll")
                if is_vol and not volumes:
                    continue
                if is_vol:
                    shape = (rows, cols, D)
                elif name in ("disp_full", "depth"):
                    shape = (prm.rows, prm.cols)
                else:
                    shape = (rows, cols)
                out[name] = np.empty(shape, dt)
                setattr(st, name, out[name].ctypes.data)
        rc = self.lib.orc_pipeline(C.byref(cfg), _p(left), _p(right), _p(pl["map_lx"]), _p(pl["map_ly"]),
                                   _p(pl["map_rx"]), _p(pl["map_ry"]), _p(pl["a1"]), _p(pl["a2"]), _p(pl["a3"]),
                                   _p(out["out"]), C.byref(st) if stages else None)
        assert rc == 0
        return out

    def float2uint8(self, rgba: np.ndarray) -> np.ndarray:
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        h, w = rgba.shape[:2]
        out = np.empty((h, w), np.uint8)
        self.lib.orc_float2uint8(_p(rgba), _p(out), C.c_int(h), C.c_int(w))
        return out

    def pointcloud(self, depth: np.ndarray, rgba: Optional[np.ndarray], fx, fy, s, cx, cy) -> np.ndarray:
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = depth.shape
        if rgba is not None:
            rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        pc = np.empty((h * w, 6 if rgba is not None else 3), np.float32)
        self.lib.orc_pointcloud(_p(depth), _p(rgba), _p(pc), C.c_int(h), C.c_int(w), C.c_float(fx), C.c_float(fy),
                                C.c_float(s), C.c_float(cx), C.c_float(cy))
        return pc


_REF_STAGES = {"rawim0": np.uint8, "rawim1": np.uint8, "noisyim0": np.uint8, "noisyim1": np.uint8, "recim0": np.uint8, "recim1": np.uint8, "bboxim0": np.uint8,
               "bboxim1": np.uint8, "census0": np.uint32, "census1": np.uint32, "rawcost": np.uint16,
               "hsum": np.uint16, "cost": np.uint16, "L0": np.uint16, "L1": np.uint16, "L2": np.uint16,
               "LAll": np.uint16, "leftDisp": np.float32, "rightDisp": np.uint16, "filteredDisp": np.float32,
               "bboxDisp": np.float32, "depth": np.float32, "rgbDepth": np.float32}


class RefEngine:
    """The unmodified reference simsense::DepthSensorEngine (registration ctor) on the current GPU."""

    def __init__(self, prm: Params, so: str = REF_SO):
        if not os.path.exists(so):
            raise FileNotFoundError(f"{so} missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(so)
        self.lib = lib
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = ([C.c_uint32] * 4 + [C.c_float] * 4 + [C.c_uint64] + [C.c_float] * 4 + [C.c_int] * 3 +
                                   [C.c_uint32] + [C.c_int] * 7 + [C.c_void_p] * 7 + [C.c_float] * 3 + [C.c_int] +
                                   [C.c_float] * 5)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_destroy.argtypes = [C.c_void_p]
        lib.ref_compute_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_uint32] * 4
        lib.ref_compute_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_uint32] * 4
        lib.ref_get_depth.restype = C.c_long
        lib.ref_get_depth.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_get_depth_device.restype = C.c_void_p
        lib.ref_get_depth_device.argtypes = [C.c_void_p]
        lib.ref_get_point_cloud.restype = C.c_long
        lib.ref_get_point_cloud.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_get_rgb_point_cloud.restype = C.c_long
        lib.ref_get_rgb_point_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_get_rgb_point_cloud_device.restype = C.c_void_p
        lib.ref_get_rgb_point_cloud_device.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_get_stage.restype = C.c_long
        lib.ref_get_stage.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        for n in ("ref_set_penalties", "ref_set_census_window_size", "ref_set_matching_block_size"):
            getattr(lib, n).argtypes = [C.c_void_p, C.c_int, C.c_int]
        for n in ("ref_set_uniqueness_ratio", "ref_set_lr_max_diff"):
            getattr(lib, n).argtypes = [C.c_void_p, C.c_int]
        self.prm = prm
        self._planes = prm.planes()  # keep alive
        pl = self._planes
        self.h = lib.ref_create(prm.rows, prm.cols, prm.rgb_rows, prm.rgb_cols, prm.focal_len, prm.baseline_len,
                                prm.min_depth, prm.max_depth, prm.ir_noise_seed, prm.speckle_shape, prm.speckle_scale,
                                prm.gaussian_mu, prm.gaussian_sigma, int(prm.rectified), prm.census_width,
                                prm.census_height, prm.max_disp, prm.bf_width, prm.bf_height, prm.p1, prm.p2,
                                prm.uniq_ratio, prm.lr_max_diff & 0xFF, prm.mf_size, _p(pl["map_lx"]), _p(pl["map_ly"]),
                                _p(pl["map_rx"]), _p(pl["map_ry"]), _p(pl["a1"]), _p(pl["a2"]), _p(pl["a3"]),
                                prm.b1, prm.b2, prm.b3, int(prm.dilation), prm.main_fx, prm.main_fy, prm.main_skew,
                                prm.main_cx, prm.main_cy)
        if not self.h:
            raise RuntimeError(lib.ref_last_error().decode())
        self.matched = prm.rows * prm.cols
        self.mshape = (prm.rows, prm.cols)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _bbox(self, bbox):
        if bbox is None:
            self.matched, self.mshape = self.prm.rows * self.prm.cols, (self.prm.rows, self.prm.cols)
            return (0, 0, 0, 0, 0)
        self.matched, self.mshape = bbox[2] * bbox[3], (bbox[3], bbox[2])
        return (1, *bbox)

    def compute_host(self, left: np.ndarray, right: np.ndarray, bbox=None):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        if self.lib.ref_compute_host(self.h, _p(left), _p(right), *self._bbox(bbox)):
            raise RuntimeError(self.lib.ref_last_error().decode())

    def compute_device(self, left_ptr: int, right_ptr: int, bbox=None):
        if self.lib.ref_compute_device(self.h, left_ptr, right_ptr, *self._bbox(bbox)):
            raise RuntimeError(self.lib.ref_last_error().decode())

    def depth(self) -> np.ndarray:
        out = np.empty((self.prm.rgb_rows, self.prm.rgb_cols), np.float32)
        if self.lib.ref_get_depth(self.h, _p(out)) < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out

    def point_cloud(self) -> np.ndarray:
        out = np.empty((self.prm.rgb_rows * self.prm.rgb_cols, 3), np.float32)
        if self.lib.ref_get_point_cloud(self.h, _p(out)) < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out

    def rgb_point_cloud(self, rgba_ptr: int) -> np.ndarray:
        out = np.empty((self.prm.rgb_rows * self.prm.rgb_cols, 6), np.float32)
        if self.lib.ref_get_rgb_point_cloud(self.h, rgba_ptr, _p(out)) < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out

    def rgb_point_cloud_device(self, rgba_ptr: int) -> int:
        p = self.lib.ref_get_rgb_point_cloud_device(self.h, rgba_ptr)
        if not p:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return p

    def stage(self, name: str) -> np.ndarray:
        dt = np.dtype(_REF_STAGES[name])
        D = self.prm.max_disp
        fsz = self.prm.rows * self.prm.cols
        cap = max(self.matched * D * 2, fsz * 4, self.prm.rgb_rows * self.prm.rgb_cols * 4)
        buf = np.empty(cap, np.uint8)
        n = self.lib.ref_get_stage(self.h, name.encode(), self.matched, _p(buf), cap)
        if n < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        a = buf[:n].view(dt)
        if name in ("rawcost", "hsum", "cost", "L0", "L1", "L2", "LAll"):
            return a.reshape(*self.mshape, D)
        if name in ("rawim0", "rawim1", "noisyim0", "noisyim1", "recim0", "recim1", "bboxDisp", "depth"):
            return a.reshape(self.prm.rows, self.prm.cols)
        if name == "rgbDepth":
            return a.reshape(self.prm.rgb_rows, self.prm.rgb_cols)
        return a.reshape(self.mshape)
