// C-ABI entry points of the SO3 / so3 op family (fp32 + fp64); see include/b200pose.h.
#include "lie_kernels.cuh"
namespace b200pose {
B200_GROUP_OPS(so3, SO3, SO3g, float, f32)
B200_GROUP_OPS(so3, SO3, SO3g, double, f64)
}  // namespace b200pose
namespace b200pose {
template <class G, typename T> using OpSo3JrG = OpSo3Jr<T>;
ABI_1_1(b200_so3_jr_f32, OpSo3JrG, SO3g, float)
ABI_1_1(b200_so3_jr_f64, OpSo3JrG, SO3g, double)
}  // namespace b200pose
