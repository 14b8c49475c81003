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

// ---- LM per-block math (csrc/lm_math.cuh) on the host, same purpose as above -------------------------------
#include "lm_math.cuh"

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
