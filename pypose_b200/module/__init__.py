from .reproj import PoseReproj
from .imu_preintegrator import IMUPreintegrator
