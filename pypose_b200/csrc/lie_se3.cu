// C-ABI entry points of the SE3 / se3 op family (fp32 + fp64); see include/b200pose.h.
#include "lie_kernels.cuh"
namespace b200pose {
B200_GROUP_OPS(se3, SE3, SE3g, float, f32)
B200_GROUP_OPS(se3, SE3, SE3g, double, f64)
}  // namespace b200pose
