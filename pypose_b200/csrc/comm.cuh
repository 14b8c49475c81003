// comm.cuh — NVLink peer-memory exchange used by the multi-GPU LM kernels (one process per GPU).
//
// Every rank owns one device buffer (cudaMalloc, exported with cudaIpcGetMemHandle and opened by the peers through
// cudaIpcOpenMemHandle; the handles travel through torch.distributed — plumbing only).  Kernels receive the table of the
// N base pointers by value and move data with plain stores over NVLink / NVSwitch: a producer kernel writes its partial
// results straight into the consumer's memory, publishes an epoch number in the consumer's flag word and the consumer
// kernel spins on its own (local) flags.  There is no collective call between the kernels of an LM step.
//
// Buffer layout (bytes):   [0, kFlagBytes)            flags: channel c, source rank r at ((c * kMaxRanks + r) * 16) bytes
//                          [kFlagBytes, kFlagBytes + kScalarBytes)   scalar slots: channel c, source rank r, 8 doubles
//                          [kDataOffset, ...)         payload regions laid out by the caller
// Memory ordering: producers issue their remote stores, `__threadfence_system()`, then take a device-scope ticket; the
// last producer CTA fences again and writes the flags (st.release.sys).  Consumers read flags with ld.acquire.sys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200pose {

constexpr int kMaxRanks = 8;
constexpr int kCommChannels = 8;
constexpr int kFlagBytes = kCommChannels * kMaxRanks * 16;
constexpr int kScalarDoubles = 8;
constexpr int kScalarBytes = kCommChannels * kMaxRanks * kScalarDoubles * 8;
constexpr int kDataOffset = 16384;       // payload starts here (flags + scalars + slack, 16 KiB aligned)

struct Peers {
  char* base[kMaxRanks];                 // base[r]: rank r's buffer as mapped into THIS process (base[rank] is local)
  int rank, world;
};

__device__ __forceinline__ unsigned long long* comm_flag(char* base, int channel, int src) {
  return reinterpret_cast<unsigned long long*>(base + (channel * kMaxRanks + src) * 16);
}
__device__ __forceinline__ double* comm_scalars(char* base, int channel, int src) {
  return reinterpret_cast<double*>(base + kFlagBytes + (channel * kMaxRanks + src) * kScalarDoubles * 8);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by ONE thread after all of this rank's payload stores are ordered before it (ticket pattern, see above):
// publish `epoch` on `channel` in every rank's buffer (including our own).
__device__ __forceinline__ void comm_signal_all(const Peers& P, int channel, unsigned long long epoch) {
  __threadfence_system();
  for (int r = 0; r < P.world; ++r) st_release_sys(comm_flag(P.base[r], channel, P.rank), epoch);
}
// publish to one rank only
__device__ __forceinline__ void comm_signal_one(const Peers& P, int dst, int channel, unsigned long long epoch) {
  __threadfence_system();
  st_release_sys(comm_flag(P.base[dst], channel, P.rank), epoch);
}
// Spin until every rank has published >= epoch on `channel` in OUR buffer.  Called by one thread per CTA, followed by
// __syncthreads(); the producers run on other GPUs, so spinning cannot starve them.
__device__ __forceinline__ void comm_wait_all(const Peers& P, int channel, unsigned long long epoch) {
  for (int r = 0; r < P.world; ++r) {
    const unsigned long long* f = comm_flag(P.base[P.rank], channel, r);
    while (ld_acquire_sys(f) < epoch) __nanosleep(20);
  }
}

// comm.cu: out = sum over ranks of in, three kernels on `s` (scatter to slice owners, reduce in rank order, gather);
// kernels return immediately when cg != NULL and cg[5] != 0 (a finished CG solve)
template <typename T>
int comm_allreduce_launch(const T* in, T* out, long long n, const Peers& P, long long stage, long long result, int channel,
                          unsigned long long epoch, unsigned* tickets, const double* cg, cudaStream_t s);

}  // namespace b200pose
