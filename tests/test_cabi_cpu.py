"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ss_b200.h declares, validates configurations before touching a device, and refuses to run
without a GPU (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ss_b200.h")
LIB = os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from sapien_b200 import _build

        _build.build_lib()
    lib = C.CDLL(LIB)
    lib.ss_last_error.restype = C.c_char_p
    lib.ss_version.restype = C.c_char_p
    return lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


class SsConfig(C.Structure):
    _fields_ = [("rows", C.c_uint32), ("cols", C.c_uint32), ("rgb_rows", C.c_uint32), ("rgb_cols", C.c_uint32),
                ("focal_len", C.c_float), ("baseline_len", C.c_float), ("min_depth", C.c_float), ("max_depth", C.c_float),
                ("ir_noise_seed", C.c_uint64), ("speckle_shape", C.c_float), ("speckle_scale", C.c_float),
                ("gaussian_mu", C.c_float), ("gaussian_sigma", C.c_float), ("rectified", C.c_int32),
                ("census_width", C.c_int32), ("census_height", C.c_int32), ("max_disp", C.c_int32),
                ("bf_width", C.c_int32), ("bf_height", C.c_int32), ("p1", C.c_int32), ("p2", C.c_int32),
                ("uniq_ratio", C.c_int32), ("lr_max_diff", C.c_int32), ("mf_size", C.c_int32),
                ("b1", C.c_float), ("b2", C.c_float), ("b3", C.c_float), ("dilation", C.c_int32),
                ("main_fx", C.c_float), ("main_fy", C.c_float), ("main_skew", C.c_float), ("main_cx", C.c_float),
                ("main_cy", C.c_float), ("registration", C.c_int32), ("device", C.c_int32), ("batch", C.c_int32),
                ("keep_stages", C.c_int32), ("lanes", C.c_int32)]


def good_config(**over):
    c = SsConfig(rows=64, cols=96, rgb_rows=96, rgb_cols=144, focal_len=100.0, baseline_len=0.05, min_depth=0.2,
                 max_depth=10.0, ir_noise_seed=0, rectified=1, census_width=7, census_height=7, max_disp=32,
                 bf_width=7, bf_height=7, p1=8, p2=32, uniq_ratio=15, lr_max_diff=1, mf_size=3, dilation=1,
                 main_fx=1, main_fy=1, registration=0, device=0, batch=1, keep_stages=0)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("ss_create", "ss_destroy", "ss_compute_host_u8", "ss_compute_device_rgba_f32", "ss_compute_device_u8",
                 "ss_get_depth_host", "ss_bind_output_host", "ss_get_stream", "ss_get_depth_device", "ss_get_point_cloud_host", "ss_get_point_cloud_device",
                 "ss_get_rgb_point_cloud_host", "ss_get_rgb_point_cloud_device", "ss_set_ir_noise_parameters",
                 "ss_set_penalties", "ss_set_census_window_size", "ss_set_matching_block_size",
                 "ss_set_uniqueness_ratio", "ss_set_lr_max_diff", "ss_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(lib):
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in include/ss_b200.h but not exported by libss_b200.so: {missing}"


def test_no_torch_or_python_in_the_abi_library():
    import subprocess

    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out, out


def test_version_string(lib):
    assert b"sm_100a" in lib.ss_version()


@pytest.mark.parametrize("over,fragment", [
    (dict(rows=16), "no less than 32"),
    (dict(census_width=4), "census_width"),
    (dict(census_width=9, census_height=9), "census_width"),
    (dict(max_disp=16), "max_disp"),
    (dict(max_disp=2048), "max_disp"),
    (dict(bf_width=2), "block_width"),
    (dict(bf_width=17, bf_height=17), "block_width"),
    (dict(p1=32, p2=8), "p1_penalty"),
    (dict(p2=224), "p1_penalty"),
    (dict(uniq_ratio=256), "uniqueness_ratio"),
    (dict(lr_max_diff=256), "lr_max_diff"),
    (dict(mf_size=4), "Median filter"),
    (dict(batch=0), "batch"),
])
def test_invalid_configuration_is_rejected_before_any_device_work(lib, over, fragment):
    """Same ranges and messages as python/py_package/sensor/simsense_component.py:54-134."""
    eng = C.c_void_p()
    rc = lib.ss_create(C.byref(good_config(**over)), None, None, None, None, None, None, None, C.byref(eng))
    assert rc == 1, lib.ss_last_error()  # SS_ERR_INVALID
    assert fragment in lib.ss_last_error().decode()
    assert not eng.value


def test_null_arguments(lib):
    assert lib.ss_create(None, None, None, None, None, None, None, None, None) == 1
    assert lib.ss_destroy(None) == 0
    assert lib.ss_synchronize(None) == 1
    assert lib.ss_get_depth_device(None, None) == 1


def test_no_cpu_fallback(lib):
    """Without a CUDA device ss_create must fail loudly with SS_ERR_NO_DEVICE."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    eng = C.c_void_p()
    rc = lib.ss_create(C.byref(good_config()), None, None, None, None, None, None, None, C.byref(eng))
    assert rc == 4, lib.ss_last_error()
    assert b"no CPU fallback" in lib.ss_last_error()
    assert not eng.value


def test_python_layer_raises_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oracle import configs
    from sapien_b200 import simsense

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simsense.DepthSensorEngine(*configs.params("small").engine_args())


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sapien_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sapien_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt or "simsense_ref" in txt:
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
