// DEV ONLY (not part of the package): alternative tilings of the v2 shell for the headline kernels, built into
// tools/variants/libvariants.so by tools/variants/build.sh and timed by tools/variants/ab.py.
#include "lie_kernels.cuh"
using namespace b200pose;
#define VAR(NAME, OPT, S, OS, TH, EPT)                                                          \
  extern "C" __attribute__((visibility("default"))) int NAME(const float* i0, float* o0, long long n, void* st) { \
    const float* in[1] = {i0}; float* out[1] = {o0};                                           \
    return launch_stream_tma<OPT<SE3g, float>, S, OS, TH, EPT>(in, out, n, (cudaStream_t)st);  \
  }
VAR(exp_s3o2_t256_e1, OpExpFwd, 3, 2, 256, 1)
VAR(exp_s2o2_t256_e1, OpExpFwd, 2, 2, 256, 1)
VAR(exp_s2o2_t256_e2, OpExpFwd, 2, 2, 256, 2)
VAR(exp_s3o2_t128_e2, OpExpFwd, 3, 2, 128, 2)
VAR(exp_s2o2_t512_e1, OpExpFwd, 2, 2, 512, 1)
VAR(exp_s3o2_t128_e1, OpExpFwd, 3, 2, 128, 1)
VAR(exp_s4o2_t128_e1, OpExpFwd, 4, 2, 128, 1)
VAR(exp_s3o3_t256_e1, OpExpFwd, 3, 3, 256, 1)
VAR(log_s3o2_t256_e1, OpLogFwd, 3, 2, 256, 1)
VAR(log_s2o2_t256_e2, OpLogFwd, 2, 2, 256, 2)
VAR(log_s3o2_t128_e2, OpLogFwd, 3, 2, 128, 2)
VAR(log_s3o2_t128_e1, OpLogFwd, 3, 2, 128, 1)
VAR(log_s2o2_t512_e1, OpLogFwd, 2, 2, 512, 1)
