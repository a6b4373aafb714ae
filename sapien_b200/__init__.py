"""sapien_b200 -- B200-native (sm_100a) drop-in for SAPIEN's GPU stereo depth sensor path
(``sapien.sensor.StereoDepthSensor`` over simsense's ``DepthSensorEngine``).

Only the hot path lives here: the CUDA kernels + C ABI (``csrc/``, ``include/ss_b200.h``), the
pybind module that mirrors ``sapien.pysapien.simsense`` and the host-side mirror of the reference's
``SimSenseComponent`` / ``StereoDepthSensorConfig``.  There is no CPU fallback: importing
:mod:`sapien_b200.simsense` without the built extension raises.
"""
__version__ = "0.1.0"
