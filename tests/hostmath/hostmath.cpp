// TEST-ONLY: compiles the device math headers (csrc/lie_math.cuh, lie_ops.cuh) for the host with
// g++ so that tests/test_hostmath.py can check every op functor against the oracle without a GPU.
// Never linked into or imported by the pypose_b200 package.
#include <string.h>
#include <vector>
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

// ---- LM per-block math (csrc/lm_math.cuh) on the host, same purpose as above -------------------------------
#include "lm_math.cuh"
#include "imu_cov_math.cuh"

template <typename T>
static void poseinv_trial_rows(const T* P, const T* X, T* Pt, double* sums, T scale, T dmin, T dmax, int rk, T delta, long long n) {
  sums[0] = sums[1] = sums[2] = sums[3] = 0;
  for (long long i = 0; i < n; ++i) {
    const Elem<T> Pe = load_elem<SE3g, T>(P + i * 7), Xe = load_elem<SE3g, T>(X + i * 7);
    Tang<T> r; Sys6<T> s;
    poseinv_linearize(Pe, Xe, r, s);
    T rho0, w0, rho1, w1;
    robust_eval(rk, delta, tang6_sqnorm(r), rho0, w0);
    if (rk) sys6_scale(s, w0);
    T D[6], pred;
    bool ok = sys6_damped_solve(s, scale, dmin, dmax, D, pred);
    const Elem<T> Pn = se3_retract(D, Pe);
    store_elem<SE3g, T>(Pt + i * 7, Pn);
    robust_eval(rk, delta, tang6_sqnorm(poseinv_residual(Pn, Xe)), rho1, w1);
    sums[0] += rho0; sums[1] += rho1; sums[2] += pred; sums[3] += ok ? 0 : 1;
  }
}
template <typename T>
static void pgo_linearize_rows(const T* nodes, const T* Z, const int* ei, const int* ej, T* M, T* u, double* loss, int rk, T delta, long long E) {
  *loss = 0;
  for (long long e = 0; e < E; ++e) {
    Tang<T> r; Sys6<T> s;
    pgo_linearize(load_elem<SE3g, T>(nodes + (long long)ei[e] * 7), load_elem<SE3g, T>(nodes + (long long)ej[e] * 7),
                  load_elem<SE3g, T>(Z + e * 7), r, s);
    T rho, w;
    robust_eval(rk, delta, tang6_sqnorm(r), rho, w);
    if (rk) sys6_scale(s, w);
    int q = 0;
    for (int p = 0; p < 6; ++p) { u[e * 6 + p] = s.g[p]; for (int c = p; c < 6; ++c) M[e * 21 + q++] = s.A[p][c]; }
    *loss += rho;
  }
}
template <typename T>
static void reproj_rows_host(const T* poses, const T* pts, const T* pix, const int* cidx, T* r, T* J, long long m) {
  for (long long k = 0; k < m; ++k) {
    V3<T> y; T rx, ry;
    reproj_residual(load_elem<SE3g, T>(poses + (long long)cidx[k] * 7), ld3(pts + k * 3), pix[k * 2], pix[k * 2 + 1], rx, ry, y);
    T j0[6], j1[6];
    reproj_rows(y, j0, j1);
    r[k * 2] = rx; r[k * 2 + 1] = ry;
    for (int a = 0; a < 6; ++a) { J[k * 12 + a] = j0[a]; J[k * 12 + 6 + a] = j1[a]; }
  }
}
extern "C" void hostmath_poseinv_trial(int is64, const void* P, const void* X, void* Pt, double* sums, double scale, double dmin,
                                       double dmax, int rk, double delta, long long n) {
  if (is64) poseinv_trial_rows<double>((const double*)P, (const double*)X, (double*)Pt, sums, scale, dmin, dmax, rk, delta, n);
  else poseinv_trial_rows<float>((const float*)P, (const float*)X, (float*)Pt, sums, (float)scale, (float)dmin, (float)dmax, rk, (float)delta, n);
}
extern "C" void hostmath_pgo_linearize(int is64, const void* nodes, const void* Z, const int* ei, const int* ej, void* M, void* u,
                                       double* loss, int rk, double delta, long long E) {
  if (is64) pgo_linearize_rows<double>((const double*)nodes, (const double*)Z, ei, ej, (double*)M, (double*)u, loss, rk, delta, E);
  else pgo_linearize_rows<float>((const float*)nodes, (const float*)Z, ei, ej, (float*)M, (float*)u, loss, rk, (float)delta, E);
}
extern "C" void hostmath_reproj_rows(int is64, const void* poses, const void* pts, const void* pix, const int* cidx, void* r, void* J, long long m) {
  if (is64) reproj_rows_host<double>((const double*)poses, (const double*)pts, (const double*)pix, cidx, (double*)r, (double*)J, m);
  else reproj_rows_host<float>((const float*)poses, (const float*)pts, (const float*)pix, cidx, (float*)r, (float*)J, m);
}

// ---- packed SPD block inverses (lm_math.cuh spd_inverse; used by the device PCG for 6x6 / 3x3 blocks)
template <typename T> static void spd_inverse_rows(int k, const T* A, T* Ai, long long n) {
  for (long long i = 0; i < n; ++i) {
    if (k == 6) {
      T M[6][6], R[6][6];
      sym6_unpack(A + i * 21, M);
      spd_inverse<T, 6>(M, R);
      sym6_pack(R, Ai + i * 21);
    } else {
      T M[3][3], R[3][3];
      sym3_unpack(A + i * 6, M);
      spd_inverse<T, 3>(M, R);
      T* o = Ai + i * 6;
      o[0] = R[0][0]; o[1] = R[0][1]; o[2] = R[0][2]; o[3] = R[1][1]; o[4] = R[1][2]; o[5] = R[2][2];
    }
  }
}
extern "C" void hostmath_spd_inverse(int is64, int k, const void* A, void* Ai, long long n) {
  if (is64) spd_inverse_rows<double>(k, (const double*)A, (double*)Ai, n);
  else spd_inverse_rows<float>(k, (const float*)A, (float*)Ai, n);
}
// ---- pose-graph linearisation with information matrices (lm_math.cuh pgo_linearize_w)
template <typename T>
static void pgo_linearize_w_rows(const T* nodes, const T* Z, const int* ei, const int* ej, const T* W, long long w_stride, T* M, T* u,
                                 T* M0, T* u0, long long E) {
  for (long long e = 0; e < E; ++e) {
    Tang<T> r; Sys6<T> sw, s0;
    pgo_linearize_w(load_elem<SE3g, T>(nodes + (long long)ei[e] * 7), load_elem<SE3g, T>(nodes + (long long)ej[e] * 7),
                    load_elem<SE3g, T>(Z + e * 7), W + e * w_stride, r, sw, s0);
    int q = 0;
    for (int p = 0; p < 6; ++p) {
      u[e * 6 + p] = sw.g[p]; u0[e * 6 + p] = s0.g[p];
      for (int c = p; c < 6; ++c) { M[e * 21 + q] = sw.A[p][c]; M0[e * 21 + q] = s0.A[p][c]; ++q; }
    }
  }
}
extern "C" void hostmath_pgo_linearize_w(int is64, const void* nodes, const void* Z, const int* ei, const int* ej, const void* W,
                                         long long w_stride, void* M, void* u, void* M0, void* u0, long long E) {
  if (is64) pgo_linearize_w_rows<double>((const double*)nodes, (const double*)Z, ei, ej, (const double*)W, w_stride, (double*)M, (double*)u, (double*)M0, (double*)u0, E);
  else pgo_linearize_w_rows<float>((const float*)nodes, (const float*)Z, ei, ej, (const float*)W, w_stride, (float*)M, (float*)u, (float*)M0, (float*)u0, E);
}
// ---- bundle-adjustment rows rebuilt from y = T p and the camera quaternion (pcg.cu obs_rows)
template <typename T>
static void ba_rows_host(const T* poses, const T* points, const int* cidx, const int* pidx, T* Jc, T* Jp, long long m) {
  for (long long k = 0; k < m; ++k) {
    const Elem<T> Tc = load_elem<SE3g, T>(poses + (long long)cidx[k] * 7);
    const V3<T> y = g_act<SE3g, T>(Tc, ld3(points + (long long)pidx[k] * 3));
    Elem<T> Q; Q.q = Tc.q;                       // obs_rows only has the quaternion
    T j0[6], j1[6], p0[3], p1[3];
    reproj_rows(y, j0, j1);
    reproj_point_rows(Q, y, p0, p1);
    for (int a = 0; a < 6; ++a) { Jc[k * 12 + a] = j0[a]; Jc[k * 12 + 6 + a] = j1[a]; }
    for (int a = 0; a < 3; ++a) { Jp[k * 6 + a] = p0[a]; Jp[k * 6 + 3 + a] = p1[a]; }
  }
}
extern "C" void hostmath_ba_rows(int is64, const void* poses, const void* points, const int* cidx, const int* pidx, void* Jc, void* Jp, long long m) {
  if (is64) ba_rows_host<double>((const double*)poses, (const double*)points, cidx, pidx, (double*)Jc, (double*)Jp, m);
  else ba_rows_host<float>((const float*)poses, (const float*)points, cidx, pidx, (float*)Jc, (float*)Jp, m);
}
// ---- IMU covariance: the three-pass chunked algorithm of scan.cu with the structured algebra of imu_cov_math.cuh,
// executed sequentially (same functions, same order of operations per chunk)
template <typename T>
static void imu_cov_host(const T* Rk, const T* Rij, const T* a, const T* dt, const T* gcov, const T* acov, long long cov_stride_f,
                         const T* init_cov, T* cov, long long F, long long chunk) {
  const long long NC = (F + chunk - 1) / chunk;
  std::vector<CovL<T> > P(NC), S(NC + 1);
  for (long long c = 0; c < NC; ++c) {
    CovL<T> L; covl_identity(L);
    const long long lo = c * chunk, hi = (lo + chunk < F) ? lo + chunk : F;
    for (long long j = hi - 1; j >= lo; --j) {
      T R[3][3];
      quat_matrix(ldq(Rij + j * 4), R);
      covl_apply_A(L, ldq(Rk + j * 4), R, ld3(a + j * 3), dt[j]);
    }
    P[c] = L;
  }
  covl_identity(S[NC]);
  for (long long c = NC - 1; c >= 0; --c) covl_mul(P[c], S[c + 1], S[c]);
  std::vector<T> Tsum(kCovT, T(0));
  for (long long c = 0; c < NC; ++c) {
    CovL<T> L = S[c + 1];
    T Tm[kCovT];
    for (int i = 0; i < kCovT; ++i) Tm[i] = T(0);
    const long long lo = c * chunk, hi = (lo + chunk < F) ? lo + chunk : F;
    for (long long j = hi - 1; j >= lo; --j) {
      const Q4<T> qk = ldq(Rk + j * 4);
      T R[3][3];
      quat_matrix(ldq(Rij + j * 4), R);
      covl_accum_noise(L, qk, R, dt[j], gcov + j * cov_stride_f, acov + j * cov_stride_f, Tm);
      covl_apply_A(L, qk, R, ld3(a + j * 3), dt[j]);
    }
    for (int i = 0; i < kCovT; ++i) Tsum[i] += Tm[i];
  }
  // cov = sum_c T_c + L0 init L0^T with L0 expanded to a dense 9x9
  T L0[9][9];
  for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) L0[r][c] = T(0);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { L0[r][c] = S[0].X[r][c]; L0[3 + r][c] = S[0].Y[r][c]; L0[6 + r][c] = S[0].Z[r][c]; }
  for (int r = 0; r < 3; ++r) { L0[3 + r][3 + r] = T(1); L0[6 + r][6 + r] = T(1); L0[6 + r][3 + r] = S[0].tau; }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      T v = Tsum[i <= j ? tri9(i, j) : tri9(j, i)];
      for (int k = 0; k < 9; ++k) for (int l = 0; l < 9; ++l) v += L0[i][k] * init_cov[k * 9 + l] * L0[j][l];
      cov[i * 9 + j] = v;
    }
}
extern "C" void hostmath_imu_cov(int is64, const void* Rk, const void* Rij, const void* a, const void* dt, const void* gcov, const void* acov,
                                 long long cov_stride_f, const void* init_cov, void* cov, long long F, long long chunk) {
  if (is64) imu_cov_host<double>((const double*)Rk, (const double*)Rij, (const double*)a, (const double*)dt, (const double*)gcov, (const double*)acov, cov_stride_f, (const double*)init_cov, (double*)cov, F, chunk);
  else imu_cov_host<float>((const float*)Rk, (const float*)Rij, (const float*)a, (const float*)dt, (const float*)gcov, (const float*)acov, cov_stride_f, (const float*)init_cov, (float*)cov, F, chunk);
}

// ---- LM accept / reject decision (csrc/lm_math.cuh lm_decide) -----------------------------------------------------
extern "C" void hostmath_lm_decide(const double* c, double cur, double trial, double predicted, double failed, double* st) {
  LmCtl k;
  k.last = c[0]; k.cached = c[1] != 0.0; k.damping = c[2]; k.pg_down = c[3]; k.reject_count = c[4]; k.reject_limit = c[5];
  k.kind = (int)c[6]; k.high = c[7]; k.low = c[8]; k.up = c[9]; k.self_down = c[10]; k.factor = c[11]; k.smin = c[12];
  k.smax = c[13];
  lm_decide(k, cur, trial, predicted, failed, st);
}

// ---- two-pose reprojection rows (csrc/lm_math.cuh reproj2_*) ------------------------------------------------------
template <typename T>
static void reproj2_rows_host(const T* nodes, const T* pts, const T* pix, const int* ia, const int* ib, const double* intr,
                              T* r, T* J, long long m) {
  const Intr<T> K = {(T)intr[0], (T)intr[1], (T)intr[2], (T)intr[3], (T)intr[4]};
  for (long long k = 0; k < m; ++k) {
    const Elem<T> Ta = load_elem<SE3g, T>(nodes + (long long)ia[k] * 7), Tb = load_elem<SE3g, T>(nodes + (long long)ib[k] * 7);
    V3<T> w, y;
    reproj2_point(Ta, Tb, mk(pts[k * 3], pts[k * 3 + 1], pts[k * 3 + 2]), w, y);
    reproj2_residual(K, y, pix[k * 2], pix[k * 2 + 1], r[k * 2], r[k * 2 + 1]);
    T j0[6], j1[6];
    reproj2_rows(K, Tb, w, y, j0, j1);
    for (int q = 0; q < 6; ++q) { J[k * 12 + q] = j0[q]; J[k * 12 + 6 + q] = j1[q]; }
  }
}
extern "C" void hostmath_reproj2_rows(int is64, const void* nodes, const void* pts, const void* pix, const int* ia, const int* ib,
                                      const double* intr, void* r, void* J, long long m) {
  if (is64) reproj2_rows_host<double>((const double*)nodes, (const double*)pts, (const double*)pix, ia, ib, intr, (double*)r, (double*)J, m);
  else reproj2_rows_host<float>((const float*)nodes, (const float*)pts, (const float*)pix, ia, ib, intr, (float*)r, (float*)J, m);
}
