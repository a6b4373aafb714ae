"""Drop-in for ``sapien.pysapien.simsense`` (reference: python/pybind/simsense.cpp:142-179) and for
``sapien.CudaArray`` (python/pybind/sapien.cpp:271-349).

The classes are implemented in C++ (csrc/pybind.cpp) on top of the C ABI (include/ss_b200.h).
There is deliberately no Python/CPU fallback: if the native module is missing this import fails.
"""
from __future__ import annotations

try:
    from ._simsense_b200 import CudaArray, DepthSensorEngine, version  # noqa: F401
except ImportError as exc:  # pragma: no cover - exercised only on broken installs
    raise ImportError(
        "sapien_b200: the native extension is not built (run `python -c 'import __graft_entry__ as g; "
        "g.build()'` or `python sapien_b200/_build.py` at the repo root). There is no CPU fallback."
    ) from exc

__all__ = ["CudaArray", "DepthSensorEngine", "version"]
