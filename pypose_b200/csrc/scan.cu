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

// ------------------------------------------------------------------------------------------------
// IMU covariance propagation (imu_preintegrator.py:428-465) without materialising (B, F+1, 9, 9):
//   cov = sum_{k=0..F} L_k B_k L_k^T,  L_k = A_k A_{k+1} ... A_{F-1},  B_0 = init_cov, B_{j+1} = N_j (noise of sample j)
// exactly the reference's `cumprod(A.flip).flip` ordering.  Three passes over chunks of the time axis, A_j and N_j
// rebuilt on the fly from (Rk = dr_j, Rij_j, a_j, dt_j):
//   1. P_c = product of the A_j of chunk c                        (one thread per (trajectory, chunk))
//   2. S_c = P_c S_{c+1}: suffix products over chunks              (one thread per trajectory, NC steps)
//   3. T_c = sum_{j in c} L_{j+1} N_j L_{j+1}^T walking the chunk backwards from L = S_{c+1}; cov = sum_c T_c + L_0 init L_0^T
// A_j = [[Rk^T,0,0],[-Rij a^ dt, I, 0],[-Rij a^ dt^2/2, dt I, I]];  N_j = (Bg Cg Bg^T + Ba Ca Ba^T)/dt,
// Bg = [Jr(dr) dt; 0; 0], Ba = [0; Rij dt; Rij dt^2/2].
// ------------------------------------------------------------------------------------------------
template <typename T> struct CovStep {
  T Rt[3][3];   // Rk^T
  T M1[3][3];   // -Rij a^ dt
  T dt;
  T Jr[3][3];   // Jr(Log dr)
  T R[3][3];    // Rij
};
template <typename T> __device__ __forceinline__ void quat_matrix(const Q4<T>& q, T (&R)[3][3]) {
  const V3<T> c0 = qrot(q, mk(T(1), T(0), T(0))), c1 = qrot(q, mk(T(0), T(1), T(0))), c2 = qrot(q, mk(T(0), T(0), T(1)));
  R[0][0] = c0.x; R[1][0] = c0.y; R[2][0] = c0.z;
  R[0][1] = c1.x; R[1][1] = c1.y; R[2][1] = c1.z;
  R[0][2] = c2.x; R[1][2] = c2.y; R[2][2] = c2.z;
}
template <typename T>
__device__ __forceinline__ void cov_step_load(CovStep<T>& s, const T* Rk, const T* Rij, const T* a, const T* dt, long long j,
                                              bool need_noise) {
  const Q4<T> qk = ldq(Rk + j * 4), qi = ldq(Rij + j * 4);
  T Rkm[3][3];
  quat_matrix(qk, Rkm);
  quat_matrix(qi, s.R);
  s.dt = dt[j];
  const V3<T> av = ld3(a + j * 3);
  const T H[3][3] = {{T(0), -av.z, av.y}, {av.z, T(0), -av.x}, {-av.y, av.x, T(0)}};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.Rt[r][c] = Rkm[c][r];
      s.M1[r][c] = -(s.R[r][0] * H[0][c] + s.R[r][1] * H[1][c] + s.R[r][2] * H[2][c]) * s.dt;
    }
  if (need_noise) {
    T jc;
    const V3<T> phi = so3_log(qk, jc);
    const RotCoef<T> rc = rot_coef(phi);
    T c1 = rc.c1, c2 = rc.c2;
    if (!(rc.theta > num<T>::eps)) { c1 = T(0); c2 = T(0); }
    const T xx = phi.x * phi.x, yy = phi.y * phi.y, zz = phi.z * phi.z, xy = phi.x * phi.y, xz = phi.x * phi.z, yz = phi.y * phi.z;
    s.Jr[0][0] = T(1) - c2 * (yy + zz); s.Jr[0][1] = c1 * phi.z + c2 * xy;   s.Jr[0][2] = -c1 * phi.y + c2 * xz;
    s.Jr[1][0] = -c1 * phi.z + c2 * xy;  s.Jr[1][1] = T(1) - c2 * (xx + zz); s.Jr[1][2] = c1 * phi.x + c2 * yz;
    s.Jr[2][0] = c1 * phi.y + c2 * xz;   s.Jr[2][1] = -c1 * phi.x + c2 * yz;  s.Jr[2][2] = T(1) - c2 * (xx + yy);
  }
}
// L <- A_j L  (L is 9x9 row-major in local memory)
template <typename T> __device__ __forceinline__ void cov_apply_A(const CovStep<T>& s, T* L) {
  const T hdt = T(0.5) * s.dt;
#pragma unroll 1
  for (int c = 0; c < 9; ++c) {
    const T l0 = L[0 * 9 + c], l1 = L[1 * 9 + c], l2 = L[2 * 9 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const T m = s.M1[r][0] * l0 + s.M1[r][1] * l1 + s.M1[r][2] * l2;
      const T mid = L[(3 + r) * 9 + c];
      L[(6 + r) * 9 + c] += hdt * m + s.dt * mid;      // M2 = M1 dt / 2
      L[(3 + r) * 9 + c] = mid + m;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) L[r * 9 + c] = s.Rt[r][0] * l0 + s.Rt[r][1] * l1 + s.Rt[r][2] * l2;
  }
}
// Tm += L N_j L^T,  N_j = (Bg Cg Bg^T + Ba Ca Ba^T)/dt
template <typename T>
__device__ __forceinline__ void cov_accum_noise(const CovStep<T>& s, const T* L, T* Tm, const T* cg, const T* ca) {
  const T hdt = T(0.5) * s.dt;
  T U[9][6];     // columns 0-2: L Bg, 3-5: L Ba
#pragma unroll 1
  for (int i = 0; i < 9; ++i) {
    const T g0 = L[i * 9 + 0], g1 = L[i * 9 + 1], g2 = L[i * 9 + 2];
    const T b0 = L[i * 9 + 3] + hdt * L[i * 9 + 6], b1 = L[i * 9 + 4] + hdt * L[i * 9 + 7], b2 = L[i * 9 + 5] + hdt * L[i * 9 + 8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      U[i][c] = (g0 * s.Jr[0][c] + g1 * s.Jr[1][c] + g2 * s.Jr[2][c]) * s.dt;
      U[i][3 + c] = (b0 * s.R[0][c] + b1 * s.R[1][c] + b2 * s.R[2][c]) * s.dt;
    }
  }
  const T idt = T(1) / s.dt;
  const T w[6] = {cg[0] * idt, cg[1] * idt, cg[2] * idt, ca[0] * idt, ca[1] * idt, ca[2] * idt};
#pragma unroll 1
  for (int i = 0; i < 9; ++i)
#pragma unroll 1
    for (int jj = 0; jj < 9; ++jj) {
      T acc = T(0);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc += U[i][c] * w[c] * U[jj][c];
      Tm[i * 9 + jj] += acc;
    }
}
template <typename T> __device__ __forceinline__ void mat9_identity(T* L) {
  for (int i = 0; i < 81; ++i) L[i] = (i % 10 == 0) ? T(1) : T(0);
}

template <typename T>
__global__ void imu_cov_chunk_prod_kernel(const T* Rk, const T* Rij, const T* a, const T* dt, T* P, long long F, long long chunk,
                                          long long NC, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const long long b = id / NC, c = id % NC;
  Rk += b * F * 4; Rij += b * F * 4; a += b * F * 3; dt += b * F;
  T L[81];
  mat9_identity(L);
  const long long lo = c * chunk, hi = (lo + chunk < F) ? lo + chunk : F;
  CovStep<T> s;
  for (long long j = hi - 1; j >= lo; --j) { cov_step_load(s, Rk, Rij, a, dt, j, false); cov_apply_A(s, L); }
  for (int i = 0; i < 81; ++i) P[id * 81 + i] = L[i];
}
// S[c] = P_c P_{c+1} ... P_{NC-1};  S[NC] = I     (P and S share storage layout (B, NC+1, 81); in-place backwards)
template <typename T> __global__ void imu_cov_suffix_kernel(const T* P, T* S, long long NC, long long B) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  T acc[81], tmp[81];
  mat9_identity(acc);
  for (int i = 0; i < 81; ++i) S[(b * (NC + 1) + NC) * 81 + i] = acc[i];
  for (long long c = NC - 1; c >= 0; --c) {
    const T* Pc = P + (b * NC + c) * 81;
    for (int i = 0; i < 9; ++i)
      for (int jj = 0; jj < 9; ++jj) {
        T v = T(0);
        for (int k = 0; k < 9; ++k) v += Pc[i * 9 + k] * acc[k * 9 + jj];
        tmp[i * 9 + jj] = v;
      }
    for (int i = 0; i < 81; ++i) { acc[i] = tmp[i]; S[(b * (NC + 1) + c) * 81 + i] = tmp[i]; }
  }
}
template <typename T>
__global__ void imu_cov_accum_kernel(const T* Rk, const T* Rij, const T* a, const T* dt, const T* gcov, const T* acov,
                                     long long cov_stride_b, long long cov_stride_f, const T* S, T* Tc, long long F, long long chunk,
                                     long long NC, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const long long b = id / NC, c = id % NC;
  Rk += b * F * 4; Rij += b * F * 4; a += b * F * 3; dt += b * F;
  gcov += b * cov_stride_b; acov += b * cov_stride_b;
  T L[81], Tm[81];
  for (int i = 0; i < 81; ++i) { L[i] = S[(b * (NC + 1) + c + 1) * 81 + i]; Tm[i] = T(0); }
  const long long lo = c * chunk, hi = (lo + chunk < F) ? lo + chunk : F;
  CovStep<T> s;
  for (long long j = hi - 1; j >= lo; --j) {
    cov_step_load(s, Rk, Rij, a, dt, j, true);
    cov_accum_noise(s, L, Tm, gcov + j * cov_stride_f, acov + j * cov_stride_f);   // uses L_{j+1}
    cov_apply_A(s, L);                                                              // L_j = A_j L_{j+1}
  }
  for (int i = 0; i < 81; ++i) Tc[id * 81 + i] = Tm[i];
}
// cov[b] = sum_c T_c + S_0 init_cov S_0^T
template <typename T>
__global__ void imu_cov_finish_kernel(const T* S, const T* Tc, const T* init_cov, long long init_stride, T* cov, long long NC,
                                      long long B) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const T* L0 = S + b * (NC + 1) * 81;
  const T* C0 = init_cov + b * init_stride;
  T tmp[81];
  for (int i = 0; i < 9; ++i)
    for (int jj = 0; jj < 9; ++jj) {
      T v = T(0);
      for (int k = 0; k < 9; ++k) v += L0[i * 9 + k] * C0[k * 9 + jj];
      tmp[i * 9 + jj] = v;
    }
  for (int i = 0; i < 9; ++i)
    for (int jj = 0; jj < 9; ++jj) {
      T v = T(0);
      for (int k = 0; k < 9; ++k) v += tmp[i * 9 + k] * L0[jj * 9 + k];
      for (long long c = 0; c < NC; ++c) v += Tc[(b * NC + c) * 81 + i * 9 + jj];
      cov[b * 81 + i * 9 + jj] = v;
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

// work: B * ((NC + 1) + 2 * NC) * 81 elements, NC = ceil(F / chunk)
#define IMU_COV_ABI(SFX, CT)                                                                                           \
  B200_EXPORT int b200_imu_cov_##SFX(const CT* Rk, const CT* Rij, const CT* a, const CT* dt, const CT* gyro_cov,       \
                                     const CT* acc_cov, long long cov_stride_b, long long cov_stride_f,                \
                                     const CT* init_cov, long long init_stride, CT* cov, CT* work, long long chunk,    \
                                     long long B, long long F, void* stream) {                                         \
    if (B <= 0 || F <= 0 || chunk <= 0) return 0;                                                                      \
    const long long NC = (F + chunk - 1) / chunk, total = B * NC;                                                      \
    CT* P = work;                                                                                                      \
    CT* S = P + total * 81;                                                                                            \
    CT* Tc = S + B * (NC + 1) * 81;                                                                                    \
    cudaStream_t st = (cudaStream_t)stream;                                                                            \
    imu_cov_chunk_prod_kernel<CT><<<(unsigned)((total + 63) / 64), 64, 0, st>>>(Rk, Rij, a, dt, P, F, chunk, NC, total); \
    imu_cov_suffix_kernel<CT><<<(unsigned)((B + 31) / 32), 32, 0, st>>>(P, S, NC, B);                                  \
    imu_cov_accum_kernel<CT><<<(unsigned)((total + 63) / 64), 64, 0, st>>>(Rk, Rij, a, dt, gyro_cov, acc_cov,          \
                                                                           cov_stride_b, cov_stride_f, S, Tc, F, chunk, \
                                                                           NC, total);                                 \
    imu_cov_finish_kernel<CT><<<(unsigned)((B + 31) / 32), 32, 0, st>>>(S, Tc, init_cov, init_stride, cov, NC, B);     \
    return (int)cudaGetLastError();                                                                                    \
  }
IMU_COV_ABI(f32, float)
IMU_COV_ABI(f64, double)
