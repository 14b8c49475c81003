// comm.cu — host side of the peer-memory exchange (comm.cuh) and a device-resident sum all-reduce built on it
// (C-ABI: include/b200pose.h, section COMM).  Reference counterpart: none — PyPose has no multi-GPU path; SURVEY.md §8e
// asks for "one packed all-reduce of [H | g | loss] per LM iteration" and "one all-reduce per CG iteration".
#include <string.h>
#include "comm.cuh"
#include "b200pose.h"
#include "lm_common.cuh"

namespace b200pose {

struct PeersArg { unsigned long long base[kMaxRanks]; int rank, world; };

static Peers make_peers(const unsigned long long* bases, int rank, int world) {
  Peers P;
  for (int r = 0; r < kMaxRanks; ++r) P.base[r] = r < world ? reinterpret_cast<char*>(bases[r]) : nullptr;
  P.rank = rank; P.world = world;
  return P;
}

// slice of rank k out of n elements, in units of 4 elements (16-byte aligned for float)
__host__ __device__ inline void ar_slice(long long n, int world, int k, long long& b, long long& e) {
  const long long q = ((n + 3) / 4 + world - 1) / world * 4;
  b = q * k < n ? q * k : n;
  e = b + q < n ? b + q : n;
}

// Two-shot all-reduce, step 1: every rank stores slice k of its vector into rank k's staging area [src rank][slice]
// (payload offset `stage`, N * slice_len elements); the last CTA publishes `epoch` on `channel`.
template <typename T>
__global__ void __launch_bounds__(kLmThreads) ar_scatter_kernel(const T* __restrict__ v, long long n, Peers P,
                                                                 long long stage, int channel, unsigned long long epoch,
                                                                 unsigned* ticket, const double* cg) {
  if (cg && cg[5] != 0.0) return;
  const long long q = ((n + 3) / 4 + P.world - 1) / P.world * 4;
  for (long long i = (long long)blockIdx.x * kLmThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLmThreads) {
    const int k = (int)(i / q);
    T* dst = reinterpret_cast<T*>(P.base[k] + kDataOffset + stage) + (long long)P.rank * q + (i - (long long)k * q);
    *dst = v[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      comm_signal_all(P, channel, epoch);
    }
  }
}
// step 2: the owner of slice `rank` sums the N staged copies in rank order and stores the result into every rank's
// result area (payload offset `result`, n elements); the last CTA publishes `epoch` on `channel + 1`.
template <typename T>
__global__ void __launch_bounds__(kLmThreads) ar_reduce_kernel(long long n, Peers P, long long stage, long long result,
                                                                int channel, unsigned long long epoch, unsigned* ticket,
                                                                const double* cg) {
  if (cg && cg[5] != 0.0) return;
  if (threadIdx.x == 0) comm_wait_all(P, channel, epoch);
  __syncthreads();
  long long b, e;
  ar_slice(n, P.world, P.rank, b, e);
  const long long q = ((n + 3) / 4 + P.world - 1) / P.world * 4;
  const T* st = reinterpret_cast<const T*>(P.base[P.rank] + kDataOffset + stage);
  for (long long i = b + (long long)blockIdx.x * kLmThreads + threadIdx.x; i < e; i += (long long)gridDim.x * kLmThreads) {
    T s = st[i - b];
    for (int r = 1; r < P.world; ++r) s += st[(long long)r * q + (i - b)];
    for (int r = 0; r < P.world; ++r) reinterpret_cast<T*>(P.base[r] + kDataOffset + result)[i] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      comm_signal_all(P, channel + 1, epoch);
    }
  }
}
// step 3: wait for every owner, then copy the reduced vector out of the exchange buffer
template <typename T>
__global__ void __launch_bounds__(kLmThreads) ar_gather_kernel(T* __restrict__ out, long long n, Peers P, long long result,
                                                                int channel, unsigned long long epoch, const double* cg) {
  if (cg && cg[5] != 0.0) return;
  if (threadIdx.x == 0) comm_wait_all(P, channel + 1, epoch);
  __syncthreads();
  const T* src = reinterpret_cast<const T*>(P.base[P.rank] + kDataOffset + result);
  for (long long i = (long long)blockIdx.x * kLmThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLmThreads)
    out[i] = src[i];
}

template <typename T>
int comm_allreduce_launch(const T* in, T* out, long long n, const Peers& P, long long stage, long long result, int channel,
                          unsigned long long epoch, unsigned* tickets, const double* cg, cudaStream_t s) {
  long long b, e;
  ar_slice(n, P.world, P.rank, b, e);
  ar_scatter_kernel<T><<<lm_grid(n, kLmThreads), kLmThreads, 0, s>>>(in, n, P, stage, channel, epoch, tickets, cg);
  ar_reduce_kernel<T><<<lm_grid(e - b > 0 ? e - b : 1, kLmThreads), kLmThreads, 0, s>>>(n, P, stage, result, channel, epoch,
                                                                                      tickets + 1, cg);
  ar_gather_kernel<T><<<lm_grid(n, kLmThreads), kLmThreads, 0, s>>>(out, n, P, result, channel, epoch, cg);
  return (int)cudaGetLastError();
}
template int comm_allreduce_launch<float>(const float*, float*, long long, const Peers&, long long, long long, int,
                                          unsigned long long, unsigned*, const double*, cudaStream_t);
template int comm_allreduce_launch<double>(const double*, double*, long long, const Peers&, long long, long long, int,
                                           unsigned long long, unsigned*, const double*, cudaStream_t);

}  // namespace b200pose

using namespace b200pose;

// ---- buffer management ------------------------------------------------------------------------------------------------
B200_EXPORT int b200_comm_alloc(long long bytes, void** ptr) {
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*ptr, 0, (size_t)bytes);
}
B200_EXPORT int b200_comm_free(void* ptr) { return (int)cudaFree(ptr); }
B200_EXPORT int b200_comm_export(void* ptr, void* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return (int)e;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
B200_EXPORT int b200_comm_open(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
B200_EXPORT int b200_comm_close(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
B200_EXPORT long long b200_comm_data_offset(void) { return kDataOffset; }

// out = sum over ranks of in (n elements), entirely on the device: three kernels, no collective library call.
// bases: HOST array of `world` device pointers (rank r's exchange buffer as mapped here); stage / result: byte offsets of
// the staging (world * slice elements) and result (n elements) regions inside the payload; tickets: 2 zeroed unsigneds.
#define COMM_ABI(SFX, CT)                                                                                             \
  B200_EXPORT int b200_comm_allreduce_##SFX(const CT* in, CT* out, long long n, const unsigned long long* bases,       \
                                            int rank, int world, long long stage, long long result, int channel,      \
                                            long long epoch, unsigned* tickets, void* stream) {                       \
    if (n <= 0) return 0;                                                                                             \
    const Peers P = make_peers(bases, rank, world);                                                                   \
    return comm_allreduce_launch<CT>(in, out, n, P, stage, result, channel, (unsigned long long)epoch, tickets,       \
                                     (const double*)nullptr, (cudaStream_t)stream);                                   \
  }
COMM_ABI(f32, float)
COMM_ABI(f64, double)
