// C-ABI entry points of the Sim3 / sim3 op family (fp32 + fp64); see include/b200pose.h.
#include "lie_kernels.cuh"
namespace b200pose {
B200_GROUP_OPS(sim3, Sim3, Sim3g, float, f32)
B200_GROUP_OPS(sim3, Sim3, Sim3g, double, f64)
}  // namespace b200pose
