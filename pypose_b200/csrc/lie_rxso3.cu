// C-ABI entry points of the RxSO3 / rxso3 op family (fp32 + fp64); see include/b200pose.h.
#include "lie_kernels.cuh"
namespace b200pose {
B200_GROUP_OPS(rxso3, RxSO3, RxSO3g, float, f32)
B200_GROUP_OPS(rxso3, RxSO3, RxSO3g, double, f64)
}  // namespace b200pose
