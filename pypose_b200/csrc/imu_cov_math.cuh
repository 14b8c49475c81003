// imu_cov_math.cuh — the structured 9x9 covariance-propagation algebra of scan.cu (imu_cov_* kernels), host + device so
// that tests/hostmath can check it against the dense oracle without a GPU.  See scan.cu for the derivation:
//   L = [[X,0,0],[Y,I,0],[Z,tau I,I]] (28 numbers), A_j L in two 3x3 products, L N L^T from the first block column.
#pragma once
#include "lie_math.cuh"

namespace b200pose {

template <typename T> struct CovL { T X[3][3], Y[3][3], Z[3][3], tau; };
constexpr int kCovL = 28;       // stored size of a CovL
constexpr int kCovT = 45;       // packed upper triangle of a symmetric 9x9 (row-major)

template <typename T> LM_HD void covl_identity(CovL<T>& L) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { L.X[r][c] = r == c ? T(1) : T(0); L.Y[r][c] = T(0); L.Z[r][c] = T(0); }
  L.tau = T(0);
}
template <typename T> LM_HD void covl_store(const CovL<T>& L, T* p) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { p[r * 3 + c] = L.X[r][c]; p[9 + r * 3 + c] = L.Y[r][c]; p[18 + r * 3 + c] = L.Z[r][c]; }
  p[27] = L.tau;
}
template <typename T> LM_HD void covl_load(CovL<T>& L, const T* p) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { L.X[r][c] = p[r * 3 + c]; L.Y[r][c] = p[9 + r * 3 + c]; L.Z[r][c] = p[18 + r * 3 + c]; }
  L.tau = p[27];
}
template <typename T> LM_HD void m3mul(const T (&A)[3][3], const T (&B)[3][3], T (&C)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r][c] = A[r][0] * B[0][c] + A[r][1] * B[1][c] + A[r][2] * B[2][c];
}
// R = I + 2 w [v]x + 2 [v]x^2 (same polynomial in q as qrot, lie_math.cuh)
template <typename T> LM_HD void quat_matrix(const Q4<T>& q, T (&R)[3][3]) {
  const T x2 = q.v.x + q.v.x, y2 = q.v.y + q.v.y, z2 = q.v.z + q.v.z;
  const T xx = q.v.x * x2, yy = q.v.y * y2, zz = q.v.z * z2, xy = q.v.x * y2, xz = q.v.x * z2, yz = q.v.y * z2;
  const T wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
  R[0][0] = T(1) - yy - zz; R[0][1] = xy - wz;        R[0][2] = xz + wy;
  R[1][0] = xy + wz;        R[1][1] = T(1) - xx - zz; R[1][2] = yz - wx;
  R[2][0] = xz - wy;        R[2][1] = yz + wx;        R[2][2] = T(1) - xx - yy;
}
// L <- A_j L.  R = matrix of Rij_j (returned for the noise term), qk = Rk_j.
template <typename T>
LM_HD void covl_apply_A(CovL<T>& L, const Q4<T>& qk, const T (&R)[3][3], const V3<T>& av, T dt) {
  T Rk[3][3], M1[3][3], MX[3][3], Xn[3][3];
  quat_matrix(qk, Rk);
  const T ax = av.x * dt, ay = av.y * dt, az = av.z * dt;        // M1 = -R a^ dt, column c = -R (a x e_c) dt
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    M1[r][0] = R[r][2] * ay - R[r][1] * az;
    M1[r][1] = R[r][0] * az - R[r][2] * ax;
    M1[r][2] = R[r][1] * ax - R[r][0] * ay;
  }
  m3mul(M1, L.X, MX);
  const T hdt = T(0.5) * dt;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      L.Z[r][c] += hdt * MX[r][c] + dt * L.Y[r][c];
      L.Y[r][c] += MX[r][c];
      Xn[r][c] = Rk[0][r] * L.X[0][c] + Rk[1][r] * L.X[1][c] + Rk[2][r] * L.X[2][c];      // Rk^T X
    }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) L.X[r][c] = Xn[r][c];
  L.tau += dt;
}
// C = P S for two structured matrices
template <typename T> LM_HD void covl_mul(const CovL<T>& P, const CovL<T>& S, CovL<T>& C) {
  T YX[3][3], ZX[3][3];
  m3mul(P.X, S.X, C.X);
  m3mul(P.Y, S.X, YX);
  m3mul(P.Z, S.X, ZX);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { C.Y[r][c] = YX[r][c] + S.Y[r][c]; C.Z[r][c] = ZX[r][c] + P.tau * S.Y[r][c] + S.Z[r][c]; }
  C.tau = P.tau + S.tau;
}
// packed index of (i,j), i <= j, in a row-major upper triangle of a 9x9
LM_HD constexpr int tri9(int i, int j) { return i * 9 - i * (i - 1) / 2 + (j - i); }

// Tm += L N_j L^T (upper triangle)
template <typename T>
LM_HD void covl_accum_noise(const CovL<T>& L, const Q4<T>& qk, const T (&R)[3][3], T dt, const T* cg,
                                                 const T* ca, T (&Tm)[kCovT]) {
  // Jr(Log dr)
  // Jr(Log dr) = I - c1 K + c2 K^2.  theta/2 = atan(|v|/w), so sin^2(theta/2) = |v|^2/|q|^2 and
  // sin(theta) = 2 |v| w / |q|^2 come from the quaternion itself: no sincos (the closed forms are those of rot_coef).
  T jc;
  const V3<T> phi = so3_log(qk, jc);
  const T x = dot(phi, phi), n2 = dot(qk.v, qk.v), iq = m_rcp(n2 + qk.w * qk.w);
  const T inv_th = m_rsqrt(x), th = x * inv_th, ix = inv_th * inv_th;
  const bool small = x < num<T>::small2;
  const T imag_s = T(0.5) + x * (T(-1.0 / 48) + x * (T(1.0 / 3840) + x * (T(-1.0 / 645120) + x * T(1.0 / 185794560))));
  const T c2_s = T(1.0 / 6) + x * (T(-1.0 / 120) + x * (T(1.0 / 5040) + x * (T(-1.0 / 362880) + x * T(1.0 / 39916800))));
  T c1 = small ? T(2) * imag_s * imag_s : T(2) * n2 * iq * ix;
  T c2 = small ? c2_s : (th - T(2) * m_sqrt(n2) * m_abs(qk.w) * iq) * ix * inv_th;
  if (!(x > num<T>::eps * num<T>::eps)) { c1 = T(0); c2 = T(0); }
  const T xx = phi.x * phi.x, yy = phi.y * phi.y, zz = phi.z * phi.z, xy = phi.x * phi.y, xz = phi.x * phi.z, yz = phi.y * phi.z;
  T Jr[3][3];
  Jr[0][0] = T(1) - c2 * (yy + zz); Jr[0][1] = c1 * phi.z + c2 * xy;   Jr[0][2] = -c1 * phi.y + c2 * xz;
  Jr[1][0] = -c1 * phi.z + c2 * xy;  Jr[1][1] = T(1) - c2 * (xx + zz); Jr[1][2] = c1 * phi.x + c2 * yz;
  Jr[2][0] = c1 * phi.y + c2 * xz;   Jr[2][1] = -c1 * phi.x + c2 * yz;  Jr[2][2] = T(1) - c2 * (xx + yy);
  const T wg[3] = {cg[0] * dt, cg[1] * dt, cg[2] * dt}, wa[3] = {ca[0] * dt, ca[1] * dt, ca[2] * dt};
  T Q[3][3], K[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Q[r][c] = Jr[r][0] * wg[0] * Jr[c][0] + Jr[r][1] * wg[1] * Jr[c][1] + Jr[r][2] * wg[2] * Jr[c][2];
      K[r][c] = R[r][0] * wa[0] * R[c][0] + R[r][1] * wa[1] * R[c][1] + R[r][2] * wa[2] * R[c][2];
    }
  T V[9][3];      // [X;Y;Z] Q
  m3mul(L.X, Q, *reinterpret_cast<T(*)[3][3]>(&V[0]));
  m3mul(L.Y, Q, *reinterpret_cast<T(*)[3][3]>(&V[3]));
  m3mul(L.Z, Q, *reinterpret_cast<T(*)[3][3]>(&V[6]));
  const T s = L.tau + T(0.5) * dt;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = i; j < 9; ++j) {
      const T(&Bj)[3][3] = j < 3 ? L.X : (j < 6 ? L.Y : L.Z);
      const int jr = j % 3;
      T acc = V[i][0] * Bj[jr][0] + V[i][1] * Bj[jr][1] + V[i][2] * Bj[jr][2];
      if (i >= 3) {                                            // accelerometer part: blocks (1,1), (1,2), (2,2)
        const T k = K[i % 3][jr];
        acc += (i < 6 ? (j < 6 ? k : s * k) : s * s * k);
      }
      Tm[tri9(i, j)] += acc;
    }
}


}  // namespace b200pose
