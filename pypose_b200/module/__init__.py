from .reproj import PoseReproj
