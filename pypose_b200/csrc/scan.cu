// scan.cu — single-pass group-product scans (cumprod / cummul) and the fused IMU preintegration.
//
// Reference: pypose/basics/ops.py:29-58 (`cumops_`: ceil(log2 L)+1 full passes of index_select + a ~37-op
// group multiply + index_copy_), pypose/module/imu_preintegrator.py:314-384 (`integrate`).
//
// Layout: (B, L, d) row-major, scan along L.  One CTA owns one sequence and walks it tile by tile, carrying
// the running prefix in registers/shared memory, so every element is read once and written once (the
// reference moves 2 d (log2 L + 1) words per element).  Inside a tile each thread scans CH consecutive
// elements in registers, thread totals are scanned with order-preserving warp shuffles, warp totals through
// shared memory.  The group operation is not commutative: every combine keeps (earlier, later) order.
#include <cuda_runtime.h>
#include <stdint.h>
#include "lie_math.cuh"

namespace b200pose {

#define B200_EXPORT extern "C" __attribute__((visibility("default")))
constexpr int kScanThreads = 128;

template <typename T> __device__ __forceinline__ T shfl_up(T v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
template <typename T> __device__ __forceinline__ T shfl_idx(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

template <typename T> __device__ __forceinline__ Elem<T> elem_shfl_up(const Elem<T>& e, int o) {
  Elem<T> r;
  r.t = mk(shfl_up(e.t.x, o), shfl_up(e.t.y, o), shfl_up(e.t.z, o));
  r.q.v = mk(shfl_up(e.q.v.x, o), shfl_up(e.q.v.y, o), shfl_up(e.q.v.z, o));
  r.q.w = shfl_up(e.q.w, o);
  r.s = shfl_up(e.s, o);
  return r;
}
template <typename T> __device__ __forceinline__ Elem<T> elem_identity() {
  Elem<T> e; e.t = mk(T(0), T(0), T(0)); e.q.v = mk(T(0), T(0), T(0)); e.q.w = T(1); e.s = T(1); return e;
}
// combine(earlier, later): y_i = y_{i-1} * x_i (right) or x_i * y_{i-1} (left)   (basics/ops.py:41-58)
template <class G, typename T, bool LEFT> __device__ __forceinline__ Elem<T> combine(const Elem<T>& a, const Elem<T>& b) {
  return LEFT ? g_mul<G, T>(b, a) : g_mul<G, T>(a, b);
}

// Block-wide exclusive scan of one Elem per thread (thread order = sequence order).
// Returns the exclusive prefix for this thread (identity for thread 0) and the block total via `total`.
template <class G, typename T, bool LEFT>
__device__ __forceinline__ Elem<T> block_exclusive(const Elem<T>& mine, Elem<T>& total, T* sh /* [warps+1][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kScanThreads / 32;
  Elem<T> inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Elem<T> up = elem_shfl_up(inc, o);
    if (lane >= o) inc = combine<G, T, LEFT>(up, inc);
  }
  Elem<T> excl = elem_shfl_up(inc, 1);
  if (lane == 0) excl = elem_identity<T>();
  if (lane == 31) store_elem<Sim3g, T>(sh + warp * 8, inc);      // widest layout holds every group
  __syncthreads();
  Elem<T> wpre = elem_identity<T>();
  Elem<T> tot = elem_identity<T>();
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    Elem<T> e = load_elem<Sim3g, T>(sh + w * 8);
    if (w < warp) wpre = combine<G, T, LEFT>(wpre, e);
    tot = combine<G, T, LEFT>(tot, e);
  }
  __syncthreads();
  total = tot;
  return combine<G, T, LEFT>(wpre, excl);
}

// ------------------------------------------------------------------------------------------------
// generic group cumprod
// ------------------------------------------------------------------------------------------------
template <class G, typename T, bool LEFT, int CH>
__global__ void __launch_bounds__(kScanThreads) cumprod_kernel(const T* __restrict__ in, T* __restrict__ out, long long L) {
  __shared__ T sh[(kScanThreads / 32) * 8];
  const T* src = in + (long long)blockIdx.x * L * G::D;
  T* dst = out + (long long)blockIdx.x * L * G::D;
  Elem<T> carry = elem_identity<T>();
  constexpr long long TILE = (long long)kScanThreads * CH;
  for (long long base = 0; base < L; base += TILE) {
    const long long first = base + (long long)threadIdx.x * CH;
    Elem<T> loc[CH];
    Elem<T> run = elem_identity<T>();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (first + c < L) {
        T row[G::D];
#pragma unroll
        for (int k = 0; k < G::D; ++k) row[k] = src[(first + c) * G::D + k];
        run = combine<G, T, LEFT>(run, load_elem<G, T>(row));
      }
      loc[c] = run;
    }
    Elem<T> total;
    Elem<T> pre = combine<G, T, LEFT>(carry, block_exclusive<G, T, LEFT>(run, total, sh));
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (first + c < L) {
        T row[G::D];
        store_elem<G, T>(row, combine<G, T, LEFT>(pre, loc[c]));
#pragma unroll
        for (int k = 0; k < G::D; ++k) dst[(first + c) * G::D + k] = row[k];
      }
    }
    carry = combine<G, T, LEFT>(carry, total);
  }
}

template <class G, typename T>
int launch_cumprod(const T* in, T* out, long long B, long long L, int left, cudaStream_t s) {
  if (B <= 0 || L <= 0) return 0;
  constexpr int CH = sizeof(T) == 8 ? 4 : 8;
  if (left) cumprod_kernel<G, T, true, CH><<<(unsigned)B, kScanThreads, 0, s>>>(in, out, L);
  else cumprod_kernel<G, T, false, CH><<<(unsigned)B, kScanThreads, 0, s>>>(in, out, L);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// IMU preintegration (imu_preintegrator.py:314-384), one kernel:
//   dr_k = Exp(gyro_k dt_k);  R_k = dr_1 ... dr_k (R_0 = I)                       -> Dr (= incre_r[:,1:]), w = dr
//   a_k  = acc_k - (init_rot R_k)^-1 g      (or acc_k - rot_k^-1 g when `rot` is given)   [uses R_k = after sample k]
//   dv_k = R_{k-1} a_k dt_k;   Dv_k = sum_{j<=k} dv_j                                      [rotates with R before k]
//   dp_k = Dv_{k-1} dt_k + R_{k-1} a_k dt_k^2 / 2;  Dp_k = sum dp_j;   Dt_k = sum dt_j
// The (v, p, t) recurrences form an associative scan with combine((v1,p1,t1),(v2,p2,t2)) = (v1+v2, p1+p2+v1 t2, t1+t2).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vpt { V3<T> v, p; T t; };
template <typename T> __device__ __forceinline__ Vpt<T> vpt_identity() {
  Vpt<T> e; e.v = mk(T(0), T(0), T(0)); e.p = e.v; e.t = T(0); return e;
}
template <typename T> __device__ __forceinline__ Vpt<T> vpt_combine(const Vpt<T>& a, const Vpt<T>& b) {
  Vpt<T> r; r.v = a.v + b.v; r.p = a.p + b.p + b.t * a.v; r.t = a.t + b.t; return r;
}
template <typename T> __device__ __forceinline__ Vpt<T> vpt_shfl_up(const Vpt<T>& e, int o) {
  Vpt<T> r;
  r.v = mk(shfl_up(e.v.x, o), shfl_up(e.v.y, o), shfl_up(e.v.z, o));
  r.p = mk(shfl_up(e.p.x, o), shfl_up(e.p.y, o), shfl_up(e.p.z, o));
  r.t = shfl_up(e.t, o);
  return r;
}
template <typename T>
__device__ __forceinline__ Vpt<T> vpt_block_exclusive(const Vpt<T>& mine, Vpt<T>& total, T* sh /* [warps][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kScanThreads / 32;
  Vpt<T> inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Vpt<T> up = vpt_shfl_up(inc, o);
    if (lane >= o) inc = vpt_combine(up, inc);
  }
  Vpt<T> excl = vpt_shfl_up(inc, 1);
  if (lane == 0) excl = vpt_identity<T>();
  if (lane == 31) { st3(sh + warp * 8, inc.v); st3(sh + warp * 8 + 3, inc.p); sh[warp * 8 + 6] = inc.t; }
  __syncthreads();
  Vpt<T> wpre = vpt_identity<T>(), tot = vpt_identity<T>();
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    Vpt<T> e; e.v = ld3(sh + w * 8); e.p = ld3(sh + w * 8 + 3); e.t = sh[w * 8 + 6];
    if (w < warp) wpre = vpt_combine(wpre, e);
    tot = vpt_combine(tot, e);
  }
  __syncthreads();
  total = tot;
  return vpt_combine(wpre, excl);
}

template <typename T, int CH>
__global__ void __launch_bounds__(kScanThreads) imu_integrate_kernel(
    const T* __restrict__ dt, const T* __restrict__ gyro, const T* __restrict__ acc, const T* __restrict__ rot,
    const T* __restrict__ init_rot, long long init_stride, T gx, T gy, T gz, T* __restrict__ a_out, T* __restrict__ Dp,
    T* __restrict__ Dv, T* __restrict__ Dr, T* __restrict__ Dt, T* __restrict__ w_out, const T* __restrict__ init_pos,
    const T* __restrict__ init_vel, long long pv_stride, T* __restrict__ rot_o, T* __restrict__ vel_o,
    T* __restrict__ pos_o, long long F) {
  __shared__ T sh[(kScanThreads / 32) * 8];
  const long long b = blockIdx.x;
  dt += b * F; gyro += b * F * 3; acc += b * F * 3;
  if (rot) rot += b * F * 4;
  // every output is optional: the integrate outputs (imu_preintegrator.py:383-384) and/or the predicted
  // states rot = R0 Dr, vel = v0 + R0 Dv, pos = p0 + R0 Dp + v0 Dt (imu_preintegrator.py:422-426)
  if (a_out) { a_out += b * F * 3; Dp += b * F * 3; Dv += b * F * 3; Dr += b * F * 4; Dt += b * F; w_out += b * F * 4; }
  if (rot_o) { rot_o += b * F * 4; vel_o += b * F * 3; pos_o += b * F * 3; }
  const V3<T> grav = mk(gx, gy, gz);
  Elem<T> R0 = elem_identity<T>();
  if (init_rot) R0.q = ldq(init_rot + b * init_stride);
  V3<T> p0 = mk(T(0), T(0), T(0)), v0 = p0;
  if (init_pos) p0 = ld3(init_pos + b * pv_stride);
  if (init_vel) v0 = ld3(init_vel + b * pv_stride);
  Elem<T> carryR = elem_identity<T>();
  Vpt<T> carry = vpt_identity<T>();
  constexpr long long TILE = (long long)kScanThreads * CH;
  for (long long base = 0; base < F; base += TILE) {
    const long long first = base + (long long)threadIdx.x * CH;
    // ---- rotation part: dr_k and the thread-local running product
    Elem<T> drs[CH], loc[CH];
    T dts[CH];
    Elem<T> run = elem_identity<T>();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      drs[c] = elem_identity<T>();
      dts[c] = T(0);
      if (first + c < F) {
        const long long k = first + c;
        dts[c] = dt[k];
        const V3<T> phi = dts[c] * mk(gyro[k * 3], gyro[k * 3 + 1], gyro[k * 3 + 2]);
        drs[c].q = so3_exp(phi, rot_coef(phi));
        run = g_mul<SO3g, T>(run, drs[c]);
      }
      loc[c] = run;
    }
    Elem<T> totR;
    const Elem<T> preR = g_mul<SO3g, T>(carryR, block_exclusive<SO3g, T, false>(run, totR, sh));
    // ---- translation part: a_k, dv_k, dp_k, thread-local (v, p, t) scan
    Vpt<T> vloc[CH];
    Vpt<T> vrun = vpt_identity<T>();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (first + c < F) {
        const long long k = first + c;
        const Elem<T> Rafter = g_mul<SO3g, T>(preR, loc[c]);                            // R_k
        const Elem<T> Rbefore = c == 0 ? preR : g_mul<SO3g, T>(preR, loc[c - 1]);       // R_{k-1}
        Q4<T> qg;                                                                        // rotation used for gravity
        if (rot) qg = ldq(rot + k * 4);
        else qg = qmul(R0.q, Rafter.q);
        const V3<T> ak = mk(acc[k * 3], acc[k * 3 + 1], acc[k * 3 + 2]) - qrot_t(qg, grav);
        const V3<T> Ra = qrot(Rbefore.q, ak);
        Vpt<T> e; e.v = dts[c] * Ra; e.p = (T(0.5) * dts[c] * dts[c]) * Ra; e.t = dts[c];
        vrun = vpt_combine(vrun, e);
        if (a_out) {
          st3(a_out + k * 3, ak);
          stq(Dr + k * 4, Rafter.q);
          stq(w_out + k * 4, drs[c].q);
        }
        if (rot_o) stq(rot_o + k * 4, qmul(R0.q, Rafter.q));
      }
      vloc[c] = vrun;
    }
    Vpt<T> totV;
    const Vpt<T> preV = vpt_combine(carry, vpt_block_exclusive(vrun, totV, sh));
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (first + c < F) {
        const long long k = first + c;
        const Vpt<T> y = vpt_combine(preV, vloc[c]);
        if (a_out) {
          st3(Dv + k * 3, y.v);
          st3(Dp + k * 3, y.p);
          Dt[k] = y.t;
        }
        if (rot_o) {
          st3(vel_o + k * 3, v0 + qrot(R0.q, y.v));
          st3(pos_o + k * 3, p0 + qrot(R0.q, y.p) + y.t * v0);
        }
      }
    }
    carryR = g_mul<SO3g, T>(carryR, totR);
    carry = vpt_combine(carry, totV);
  }
}

}  // namespace b200pose

using namespace b200pose;

#define SCAN_ABI(GRP, G)                                                                                               \
  B200_EXPORT int b200_##GRP##_cumprod_f32(const float* in, float* out, long long B, long long L, int left, void* s) { \
    return launch_cumprod<G, float>(in, out, B, L, left, (cudaStream_t)s);                                             \
  }                                                                                                                    \
  B200_EXPORT int b200_##GRP##_cumprod_f64(const double* in, double* out, long long B, long long L, int left, void* s) { \
    return launch_cumprod<G, double>(in, out, B, L, left, (cudaStream_t)s);                                            \
  }
SCAN_ABI(SO3, SO3g)
SCAN_ABI(SE3, SE3g)
SCAN_ABI(RxSO3, RxSO3g)
SCAN_ABI(Sim3, Sim3g)

#define IMU_ABI(SFX, CT, CH)                                                                                           \
  B200_EXPORT int b200_imu_integrate_##SFX(const CT* dt, const CT* gyro, const CT* acc, const CT* rot,                 \
                                           const CT* init_rot, long long init_stride, const CT* gravity3_host, CT* a,  \
                                           CT* Dp, CT* Dv, CT* Dr, CT* Dt, CT* w, const CT* init_pos,                  \
                                           const CT* init_vel, long long pv_stride, CT* rot_out, CT* vel_out,          \
                                           CT* pos_out, long long B, long long F, void* s) {                           \
    if (B <= 0 || F <= 0) return 0;                                                                                    \
    imu_integrate_kernel<CT, CH><<<(unsigned)B, kScanThreads, 0, (cudaStream_t)s>>>(                                   \
        dt, gyro, acc, rot, init_rot, init_stride, gravity3_host[0], gravity3_host[1], gravity3_host[2], a, Dp, Dv,    \
        Dr, Dt, w, init_pos, init_vel, pv_stride, rot_out, vel_out, pos_out, F);                                       \
    return (int)cudaGetLastError();                                                                                    \
  }
IMU_ABI(f32, float, 4)
IMU_ABI(f64, double, 2)
