// TEST-ONLY: compiles the device math headers (csrc/lie_math.cuh, lie_ops.cuh) for the host with
// g++ so that tests/test_hostmath.py can check every op functor against the oracle without a GPU.
// Never linked into or imported by the pypose_b200 package.
#include <string.h>
#include "lie_ops.cuh"

using namespace b200pose;

template <class Op>
static int run_rows(const void* const* ins, void* const* outs, long long n) {
  using T = typename Op::T;
  const T* i0 = (const T*)ins[0];
  const T* i1 = Op::NIN > 1 ? (const T*)ins[1] : nullptr;
  const T* i2 = Op::NIN > 2 ? (const T*)ins[2] : nullptr;
  T* o0 = (T*)outs[0];
  T* o1 = Op::NOUT > 1 ? (T*)outs[1] : nullptr;
  T d1[1] = {0}, d2[1] = {0}, e1[1];
  for (long long r = 0; r < n; ++r) {
    Op::apply(i0 + r * Op::DI0, i1 ? i1 + r * Op::DI1 : d1, i2 ? i2 + r * Op::DI2 : d2, o0 + r * Op::DO0,
              o1 ? o1 + r * Op::DO1 : e1);
  }
  return 0;
}

#define TRY(op, OPT, NIN_, NOUT_, ALG)                                                         \
  if (!strcmp(opname, #op)) {                                                                  \
    if (is64) return run_rows<OPT<G, double> >(ins, outs, n);                                  \
    return run_rows<OPT<G, float> >(ins, outs, n);                                             \
  }

template <class G> static int dispatch(const char* opname, int is64, const void* const* ins, void* const* outs, long long n) {
  B200_FOR_EACH_GROUP_OP(TRY)
  return -1;
}

extern "C" int hostmath_run(const char* group, const char* opname, int is64, const void* const* ins,
                            void* const* outs, long long n) {
  if (!strcmp(group, "SO3")) {
    if (!strcmp(opname, "jr")) return is64 ? run_rows<OpSo3Jr<double> >(ins, outs, n) : run_rows<OpSo3Jr<float> >(ins, outs, n);
    return dispatch<SO3g>(opname, is64, ins, outs, n);
  }
  if (!strcmp(group, "SE3")) return dispatch<SE3g>(opname, is64, ins, outs, n);
  if (!strcmp(group, "RxSO3")) return dispatch<RxSO3g>(opname, is64, ins, outs, n);
  if (!strcmp(group, "Sim3")) return dispatch<Sim3g>(opname, is64, ins, outs, n);
  return -2;
}
