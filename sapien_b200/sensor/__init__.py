from .simsense_component import SimSenseComponent  # noqa: F401
from .stereodepth import StereoDepthSensor, StereoDepthSensorConfig  # noqa: F401
