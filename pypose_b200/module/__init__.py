from .reproj import PoseReproj
from .pgo import PoseGraph
from .imu_preintegrator import IMUPreintegrator
