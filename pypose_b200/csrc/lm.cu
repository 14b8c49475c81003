// lm.cu — fused kernels of the Levenberg-Marquardt inner loop (C-ABI: include/b200pose.h, section LM).
//
// Replaces, for the residual families whose block structure is known (lm_math.cuh), the reference's
// modjac -> dense J -> J^T J -> clamp/damp -> cholesky_ex/cholesky_solve -> p.add_ -> loss sequence
// (pypose/optim/optimizer.py:645-680, optim/functional.py:130-153, optim/solver.py:213-216).
//
// Scalars that the host control flow reads (loss, trial loss, predicted reduction) are reduced on the
// device in fp64: per-CTA partials, then the last CTA to finish folds them in a fixed order, so results
// are deterministic for a given grid.  Work is enqueued on the caller's stream; nothing synchronises.
#include "lm_common.cuh"

namespace b200pose {

// ------------------------------------------------------------------------------------------------
// PoseInv (BASELINE.json configs[2], README.md:120-135 InvNet): residual Log(P_i X_i), one 6x6 system per pose.
// ------------------------------------------------------------------------------------------------
// loss = sum_i |Log(P_i X_i)|^2     (RobustModel.loss with the trivial kernel, optimizer.py:118-125)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_poseinv_loss_kernel(const T* __restrict__ P, const T* __restrict__ X,
                                                                      double* ws, int rk, T rdelta, long long n) {
  double acc[1] = {0.0};
  for (long long i = (long long)blockIdx.x * kLmThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLmThreads) {
    T p[7], x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { p[k] = P[i * 7 + k]; x[k] = X[i * 7 + k]; }
    T rho, w;
    robust_eval(rk, rdelta, tang6_sqnorm(poseinv_residual(load_se3(p), load_se3(x))), rho, w);
    acc[0] += (double)rho;
  }
  reduce_sums<1>(acc, ws);
}

// One complete LM trial per pose, entirely in registers:
//   linearise (r, J = Jl^-1(r)) -> A = J^T J, g = J^T r -> clamp/damp -> Cholesky -> D -> P' = Exp(D) P -> |r'|^2
// sums: [0] sum |r|^2 (current loss), [1] sum |r'|^2 (trial loss), [2] sum (J D)^T (2 R + J D), [3] #failed pivots
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_poseinv_trial_kernel(const T* __restrict__ P, const T* __restrict__ X,
                                                                       T* __restrict__ Pt, double* ws, T scale, T dmin,
                                                                       T dmax, int rk, T rdelta, long long n) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * kLmThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLmThreads) {
    T p[7], x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { p[k] = P[i * 7 + k]; x[k] = X[i * 7 + k]; }
    const Elem<T> Pe = load_se3(p), Xe = load_se3(x);
    Tang<T> r;
    Sys6<T> s;
    poseinv_linearize(Pe, Xe, r, s);
    T rho0, w0, rho1, w1;
    robust_eval(rk, rdelta, tang6_sqnorm(r), rho0, w0);
    if (rk) sys6_scale(s, w0);
    T D[6], pred;
    const bool ok = sys6_damped_solve(s, scale, dmin, dmax, D, pred);
    const Elem<T> Pn = se3_retract(D, Pe);
    T o[7];
    store_elem<SE3g, T>(o, Pn);
#pragma unroll
    for (int k = 0; k < 7; ++k) Pt[i * 7 + k] = o[k];
    robust_eval(rk, rdelta, tang6_sqnorm(poseinv_residual(Pn, Xe)), rho1, w1);
    acc[0] += (double)rho0;
    acc[1] += (double)rho1;
    acc[2] += (double)pred;
    acc[3] += ok ? 0.0 : 1.0;
  }
  reduce_sums<4>(acc, ws);
}

// ------------------------------------------------------------------------------------------------
// Pose-graph LM with reprojection residuals (BASELINE.json configs[4], single-pose form, SURVEY.md §8d cfg 5-min):
//   r_k = pi(T_{c_k} p_k) - z_k, observations sorted by camera, seg[c] .. seg[c+1] = camera c's rows.
// ------------------------------------------------------------------------------------------------
// One warp per camera: lanes stride over the camera's observations, accumulate the 6x6 system in registers,
// shuffle-reduce, lane 0 writes H[c] (21 upper-triangular entries, row-major) and g[c] (6): no atomics.
// sums: [0] sum |r|^2.   With residual sharding each rank sees only its rows; H/g/loss are all-reduced by the host.
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj_accum_kernel(const T* __restrict__ poses, const T* __restrict__ pts,
                                                                      const T* __restrict__ pix, const int* __restrict__ seg,
                                                                      T* __restrict__ H, T* __restrict__ g, double* ws,
                                                                      int rk, T rdelta, int ncam) {
  const int lane = threadIdx.x & 31;
  const int wpb = kLmThreads / 32;
  double acc[1] = {0.0};
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < ncam; c += gridDim.x * wpb) {
    T pr[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) pr[k] = poses[(long long)c * 7 + k];
    const Elem<T> Tc = load_se3(pr);
    Acc6<T> ac;
    ac.zero();
    T loss = T(0);
    const int b = seg[c], e = seg[c + 1];
    auto accumulate = [&](const V3<T>& p, T zx, T zy) {
      T rx, ry;
      V3<T> y;
      reproj_residual(Tc, p, zx, zy, rx, ry, y);
      T j0[6], j1[6];
      reproj_rows(y, j0, j1);
      T rho, w;
      robust_eval(rk, rdelta, rx * rx + ry * ry, rho, w);
      if (rk) {                      // FastTriggs: rows and residual scaled by sqrt(rho')
        const T sw = m_sqrt(w);
        rx *= sw; ry *= sw;
#pragma unroll
        for (int a = 0; a < 6; ++a) { j0[a] *= sw; j1[a] *= sw; }
      }
      ac.add_row(j0, rx);
      ac.add_row(j1, ry);
      loss += rho;
    };
    // two observations per lane per iteration, all ten loads issued before the math (memory-level parallelism:
    // ncu showed the one-at-a-time loop stalled on long_scoreboard with 46 % issue utilisation)
    int k = b + lane;
    for (; k + 32 < e; k += 64) {
      const long long k0 = k, k1 = k + 32;
      const V3<T> p0 = mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]);
      const V3<T> p1 = mk(pts[k1 * 3], pts[k1 * 3 + 1], pts[k1 * 3 + 2]);
      const T z0x = pix[k0 * 2], z0y = pix[k0 * 2 + 1], z1x = pix[k1 * 2], z1y = pix[k1 * 2 + 1];
      accumulate(p0, z0x, z0y);
      accumulate(p1, z1x, z1y);
    }
    if (k < e) {
      const long long k0 = k;
      accumulate(mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]), pix[k0 * 2], pix[k0 * 2 + 1]);
    }
    Sys6<T> s = ac.finish();
    // warp reduction of 21 + 6 + 1 values
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        s.g[a] += __shfl_xor_sync(0xffffffffu, s.g[a], o);
#pragma unroll
        for (int bb = a; bb < 6; ++bb) s.A[a][bb] += __shfl_xor_sync(0xffffffffu, s.A[a][bb], o);
      }
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
    }
    if (lane == 0) {
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        g[(long long)c * 6 + a] = s.g[a];
#pragma unroll
        for (int bb = a; bb < 6; ++bb) H[(long long)c * 21 + q++] = s.A[a][bb];
      }
      acc[0] += (double)loss;
    }
  }
  reduce_sums<1>(acc, ws);
}

// Generic damped batched 6x6 solve + retraction on packed blocks (also the second half of the reprojection step):
//   D_c = (clamp(diag H_c) * scale + offdiag H_c)^-1 (-g_c);  P'_c = Exp(D_c) P_c
// sums: [0] sum predicted = sum D^T H D + 2 D^T g, [1] #failed pivots
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_solve6_retract_kernel(const T* __restrict__ H, const T* __restrict__ g,
                                                                        const T* __restrict__ P, T* __restrict__ Pt,
                                                                        T* __restrict__ Dout, double* ws, T scale, T dmin,
                                                                        T dmax, long long n) {
  double acc[2] = {0.0, 0.0};
  for (long long c = (long long)blockIdx.x * kLmThreads + threadIdx.x; c < n; c += (long long)gridDim.x * kLmThreads) {
    Sys6<T> s;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      s.g[a] = g[c * 6 + a];
#pragma unroll
      for (int b = a; b < 6; ++b) s.A[a][b] = H[c * 21 + q++];
    }
    T D[6], pred;
    const bool ok = sys6_damped_solve(s, scale, dmin, dmax, D, pred);
    T pr[7], o[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) pr[k] = P[c * 7 + k];
    store_elem<SE3g, T>(o, se3_retract(D, load_se3(pr)));
#pragma unroll
    for (int k = 0; k < 7; ++k) Pt[c * 7 + k] = o[k];
    if (Dout)
#pragma unroll
      for (int k = 0; k < 6; ++k) Dout[c * 6 + k] = D[k];
    acc[0] += (double)pred;
    acc[1] += ok ? 0.0 : 1.0;
  }
  reduce_sums<2>(acc, ws);
}

// trial loss over a shard of observations: one warp per camera (pose in registers), lanes stride over the camera's
// rows two at a time — the same access pattern as the accumulate kernel without the Jacobian work.
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj_loss_kernel(const T* __restrict__ poses, const T* __restrict__ pts,
                                                                     const T* __restrict__ pix, const int* __restrict__ seg,
                                                                     double* ws, int rk, T rdelta, int ncam) {
  const int lane = threadIdx.x & 31;
  const int wpb = kLmThreads / 32;
  double acc[1] = {0.0};
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < ncam; c += gridDim.x * wpb) {
    T pr[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) pr[q] = poses[(long long)c * 7 + q];
    const Elem<T> Tc = load_se3(pr);
    const int b = seg[c], e = seg[c + 1];
    T loss = T(0);
    auto one = [&](const V3<T>& p, T zx, T zy) {
      T rx, ry, rho, w;
      V3<T> y;
      reproj_residual(Tc, p, zx, zy, rx, ry, y);
      robust_eval(rk, rdelta, rx * rx + ry * ry, rho, w);
      loss += rho;
    };
    int k = b + lane;
    for (; k + 32 < e; k += 64) {
      const long long k0 = k, k1 = k + 32;
      const V3<T> p0 = mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]);
      const V3<T> p1 = mk(pts[k1 * 3], pts[k1 * 3 + 1], pts[k1 * 3 + 2]);
      const T z0x = pix[k0 * 2], z0y = pix[k0 * 2 + 1], z1x = pix[k1 * 2], z1y = pix[k1 * 2 + 1];
      one(p0, z0x, z0y);
      one(p1, z1x, z1y);
    }
    if (k < e) {
      const long long k0 = k;
      one(mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]), pix[k0 * 2], pix[k0 * 2 + 1]);
    }
    acc[0] += (double)loss;
  }
  reduce_sums<1>(acc, ws);
}

// residual vector r (m, 2) for API-level forward() parity
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj_residual_kernel(const T* __restrict__ poses, const T* __restrict__ pts,
                                                                         const T* __restrict__ pix, const int* __restrict__ cidx,
                                                                         T* __restrict__ r, long long m) {
  for (long long k = (long long)blockIdx.x * kLmThreads + threadIdx.x; k < m; k += (long long)gridDim.x * kLmThreads) {
    const int c = cidx[k];
    T pr[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) pr[q] = __ldg(poses + (long long)c * 7 + q);
    const V3<T> p = mk(pts[k * 3], pts[k * 3 + 1], pts[k * 3 + 2]);
    T rx, ry;
    V3<T> y;
    reproj_residual(load_se3(pr), p, pix[k * 2], pix[k * 2 + 1], rx, ry, y);
    r[k * 2] = rx;
    r[k * 2 + 1] = ry;
  }
}

// ------------------------------------------------------------------------------------------------
// Pose-graph optimisation, two-pose residuals r_e = Log(Z_e^-1 A^-1 B), A = nodes[ei], B = nodes[ej]
// (examples/module/pgo/pgo.py:15-25).  H is block-sparse: the edge's M_e = J^T J goes to (i,i), (j,j) and -M_e to
// (i,j), (j,i).  M_e (21) and u_e = J^T r (6) are stored per edge; H is never assembled — the PCG multiplies with it
// edge by edge (what the reference delegates to the external `bae` package, optimizer.py:629-643).
// ------------------------------------------------------------------------------------------------
// One edge's blocks, per edge (M, u: the scatter / multi-GPU route) and / or in node order (Mn, un: 24-float slots at the
// edge's position in its first node's list — the node whose Jacobian is -J — and in its second node's list; pcg2.cu)
template <typename T>
__device__ __forceinline__ void store_edge_blocks(const Sys6<T>& s, long long e, T* __restrict__ M, T* __restrict__ u,
                                                  const int* __restrict__ epos_i, const int* __restrict__ epos_j,
                                                  T* __restrict__ Mn, T* __restrict__ un) {
  if (M) {
    int q = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      u[e * 6 + p] = s.g[p];
#pragma unroll
      for (int c = p; c < 6; ++c) M[e * 21 + q++] = s.A[p][c];
    }
  }
  if (Mn) {
    const long long si = epos_i[e], sj = epos_j[e];
    int q = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      un[si * 6 + p] = -s.g[p];
      un[sj * 6 + p] = s.g[p];
#pragma unroll
      for (int c = p; c < 6; ++c) { Mn[si * 24 + q] = s.A[p][c]; Mn[sj * 24 + q] = s.A[p][c]; ++q; }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_pgo_linearize_kernel(const T* __restrict__ nodes, const T* __restrict__ Z,
                                                                       const int* __restrict__ ei, const int* __restrict__ ej,
                                                                       T* __restrict__ M, T* __restrict__ u,
                                                                       const int* __restrict__ epos_i,
                                                                       const int* __restrict__ epos_j, T* __restrict__ Mn,
                                                                       T* __restrict__ un, double* ws, int rk,
                                                                       T rdelta, long long E) {
  double acc[1] = {0.0};
  for (long long e = (long long)blockIdx.x * kLmThreads + threadIdx.x; e < E; e += (long long)gridDim.x * kLmThreads) {
    T a[7], b[7], z[7];
    const long long i = ei[e], j = ej[e];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = __ldg(nodes + i * 7 + k); b[k] = __ldg(nodes + j * 7 + k); z[k] = Z[e * 7 + k]; }
    Tang<T> r;
    Sys6<T> s;
    pgo_linearize(load_se3(a), load_se3(b), load_se3(z), r, s);
    T rho, w;
    robust_eval(rk, rdelta, tang6_sqnorm(r), rho, w);
    if (rk) sys6_scale(s, w);
    store_edge_blocks(s, e, M, u, epos_i, epos_j, Mn, un);
    acc[0] += (double)rho;
  }
  reduce_sums<1>(acc, ws);
}

// The same with per-edge information matrices W_e (examples/module/pgo/pgo.py:75 `weight=infos`): M = w J^T W J,
// u = w J^T W r for the solve, M0 = w J^T J and u0 = w J^T r for the step-quality term (w = robust weight).
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_pgo_linearize_w_kernel(const T* __restrict__ nodes, const T* __restrict__ Z,
                                                                         const int* __restrict__ ei, const int* __restrict__ ej,
                                                                         const T* __restrict__ W, long long w_stride,
                                                                         T* __restrict__ M, T* __restrict__ u, T* __restrict__ M0,
                                                                         T* __restrict__ u0, double* ws, int rk, T rdelta,
                                                                         long long E) {
  double acc[1] = {0.0};
  for (long long e = (long long)blockIdx.x * kLmThreads + threadIdx.x; e < E; e += (long long)gridDim.x * kLmThreads) {
    T a[7], b[7], z[7], w36[36];
    const long long i = ei[e], j = ej[e];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = __ldg(nodes + i * 7 + k); b[k] = __ldg(nodes + j * 7 + k); z[k] = Z[e * 7 + k]; }
#pragma unroll
    for (int k = 0; k < 36; ++k) w36[k] = W[e * w_stride + k];
    Tang<T> r;
    Sys6<T> sw, s0;
    pgo_linearize_w(load_se3(a), load_se3(b), load_se3(z), w36, r, sw, s0);
    T rho, w;
    robust_eval(rk, rdelta, tang6_sqnorm(r), rho, w);
    if (rk) { sys6_scale(sw, w); sys6_scale(s0, w); }
    int q = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      u[e * 6 + p] = sw.g[p];
      u0[e * 6 + p] = s0.g[p];
#pragma unroll
      for (int c = p; c < 6; ++c) { M[e * 21 + q] = sw.A[p][c]; M0[e * 21 + q] = s0.A[p][c]; ++q; }
    }
    acc[0] += (double)rho;
  }
  reduce_sums<1>(acc, ws);
}

// Hd[i] += M_e, Hd[j] += M_e (diagonal blocks, packed 21);  g[i] -= u_e, g[j] += u_e   (J_A = -J, J_B = +J)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_pgo_scatter_kernel(const T* __restrict__ M, const T* __restrict__ u,
                                                                     const int* __restrict__ ei, const int* __restrict__ ej,
                                                                     T* __restrict__ Hd, T* __restrict__ g, long long E) {
  for (long long e = (long long)blockIdx.x * kLmThreads + threadIdx.x; e < E; e += (long long)gridDim.x * kLmThreads) {
    const long long i = ei[e], j = ej[e];
#pragma unroll
    for (int k = 0; k < 21; ++k) { const T m = M[e * 21 + k]; atomicAdd(Hd + i * 21 + k, m); atomicAdd(Hd + j * 21 + k, m); }
#pragma unroll
    for (int k = 0; k < 6; ++k) { const T v = u[e * 6 + k]; atomicAdd(g + i * 6 + k, -v); atomicAdd(g + j * 6 + k, v); }
  }
}

// y += H x edge by edge: v = M_e (x_i - x_j); y_i += v; y_j -= v      (y pre-initialised by the caller)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_pgo_spmv_kernel(const T* __restrict__ M, const int* __restrict__ ei,
                                                                  const int* __restrict__ ej, const T* __restrict__ x,
                                                                  T* __restrict__ y, long long E) {
  for (long long e = (long long)blockIdx.x * kLmThreads + threadIdx.x; e < E; e += (long long)gridDim.x * kLmThreads) {
    const long long i = ei[e], j = ej[e];
    T d[6], A[6][6];
    int q = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      d[p] = __ldg(x + i * 6 + p) - __ldg(x + j * 6 + p);
#pragma unroll
      for (int c = p; c < 6; ++c) { A[p][c] = M[e * 21 + q]; A[c][p] = A[p][c]; ++q; }
    }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      T v = T(0);
#pragma unroll
      for (int c = 0; c < 6; ++c) v += A[p][c] * d[c];
      atomicAdd(y + i * 6 + p, v);
      atomicAdd(y + j * 6 + p, -v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_pgo_loss_kernel(const T* __restrict__ nodes, const T* __restrict__ Z,
                                                                  const int* __restrict__ ei, const int* __restrict__ ej,
                                                                  double* ws, int rk, T rdelta, long long E) {
  double acc[1] = {0.0};
  for (long long e = (long long)blockIdx.x * kLmThreads + threadIdx.x; e < E; e += (long long)gridDim.x * kLmThreads) {
    T a[7], b[7], z[7];
    const long long i = ei[e], j = ej[e];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = __ldg(nodes + i * 7 + k); b[k] = __ldg(nodes + j * 7 + k); z[k] = Z[e * 7 + k]; }
    Elem<T> S;
    T rho, w;
    robust_eval(rk, rdelta, tang6_sqnorm(pgo_residual(load_se3(a), load_se3(b), load_se3(z), S)), rho, w);
    acc[0] += (double)rho;
  }
  reduce_sums<1>(acc, ws);
}

// ------------------------------------------------------------------------------------------------
// Bundle adjustment: r_k = pi(T_{c_k} p_{j_k}) - z_k with BOTH the poses and the points as parameters
// (README.md:153-198 sparse example; examples/module/ba).  Per observation: Jc (2x6) and Jp (2x3), stored (already
// scaled by sqrt(rho')) together with the scaled residual so that the Schur-complement PCG can multiply with
// W = sum Jc^T Jp and W^T observation by observation without ever forming the reduced camera matrix.
// ------------------------------------------------------------------------------------------------
// Y4 (optional, (m,4)): the camera-frame point y = T p and the robust scale sqrt(rho') — 16 B from which the PCG
// kernels (pcg.cu) rebuild both row pairs instead of streaming 72 B of stored rows per observation and product.
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_ba_linearize_kernel(
    const T* __restrict__ poses, const T* __restrict__ points, const T* __restrict__ pix, const int* __restrict__ cidx,
    const int* __restrict__ pidx, T* __restrict__ Jc, T* __restrict__ Jp, T* __restrict__ Y4, const int* __restrict__ ppos,
    T* __restrict__ Y4p, T* __restrict__ rs, T* __restrict__ Hcc, T* __restrict__ Hpp, T* __restrict__ gc,
    T* __restrict__ gp, double* ws, int rk, T rdelta, long long m) {
  double acc[1] = {0.0};
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: the camera-side sums are combined across the warp before the atomics (seg_atomic_add)
  for (long long k0 = (long long)blockIdx.x * kLmThreads + (threadIdx.x - lane); k0 < m; k0 += (long long)gridDim.x * kLmThreads) {
    const long long k = k0 + lane;
    const bool active = k < m;
    long long c = 0;
    T cam[27];
#pragma unroll
    for (int a = 0; a < 27; ++a) cam[a] = T(0);
    if (active) {
      c = cidx[k];
      const long long j = pidx[k];
      T pr[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) pr[q] = __ldg(poses + c * 7 + q);
      const Elem<T> Tc = load_se3(pr);
      const V3<T> p = mk(__ldg(points + j * 3), __ldg(points + j * 3 + 1), __ldg(points + j * 3 + 2));
      T rx, ry;
      V3<T> y;
      reproj_residual(Tc, p, pix[k * 2], pix[k * 2 + 1], rx, ry, y);
      T j0[6], j1[6], p0[3], p1[3];
      reproj_rows(y, j0, j1);
      reproj_point_rows(Tc, y, p0, p1);
      T rho, w;
      robust_eval(rk, rdelta, rx * rx + ry * ry, rho, w);
      T sw = T(1);
      if (rk) {
        sw = m_sqrt(w);
        rx *= sw; ry *= sw;
#pragma unroll
        for (int a = 0; a < 6; ++a) { j0[a] *= sw; j1[a] *= sw; }
#pragma unroll
        for (int a = 0; a < 3; ++a) { p0[a] *= sw; p1[a] *= sw; }
      }
      if (Jc) {
#pragma unroll
        for (int a = 0; a < 6; ++a) { Jc[k * 12 + a] = j0[a]; Jc[k * 12 + 6 + a] = j1[a]; }
#pragma unroll
        for (int a = 0; a < 3; ++a) { Jp[k * 6 + a] = p0[a]; Jp[k * 6 + 3 + a] = p1[a]; }
      }
      if (Y4) { Y4[k * 4] = y.x; Y4[k * 4 + 1] = y.y; Y4[k * 4 + 2] = y.z; Y4[k * 4 + 3] = sw; }
      if (Y4p) {                          // the same 16 B at the observation's position in point order
        const long long s = ppos[k];
        Y4p[s * 4] = y.x; Y4p[s * 4 + 1] = y.y; Y4p[s * 4 + 2] = y.z; Y4p[s * 4 + 3] = sw;
      }
      rs[k * 2] = rx; rs[k * 2 + 1] = ry;
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        cam[21 + a] = j0[a] * rx + j1[a] * ry;
#pragma unroll
        for (int b = a; b < 6; ++b) cam[q++] = j0[a] * j0[b] + j1[a] * j1[b];
      }
      q = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        atomicAdd(gp + j * 3 + a, p0[a] * rx + p1[a] * ry);
#pragma unroll
        for (int b = a; b < 3; ++b) atomicAdd(Hpp + j * 6 + q++, p0[a] * p0[b] + p1[a] * p1[b]);
      }
      acc[0] += (double)rho;
    }
    // Hcc (21) and gc (6) live in different arrays: two segmented adds with the same key
    T hc[21], g6[6];
#pragma unroll
    for (int a = 0; a < 21; ++a) hc[a] = cam[a];
#pragma unroll
    for (int a = 0; a < 6; ++a) g6[a] = cam[21 + a];
    seg_atomic_add<T, 21>(Hcc + c * 21, c, hc, active);
    seg_atomic_add<T, 6>(gc + c * 6, c, g6, active);
  }
  reduce_sums<1>(acc, ws);
}

// t[j] += Jp_k^T (Jc_k x[c_k])      (W^T x, cameras -> points)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_ba_wtx_kernel(const T* __restrict__ Jc, const T* __restrict__ Jp,
                                                                const int* __restrict__ cidx, const int* __restrict__ pidx,
                                                                const T* __restrict__ x, T* __restrict__ t, long long m) {
  for (long long k = (long long)blockIdx.x * kLmThreads + threadIdx.x; k < m; k += (long long)gridDim.x * kLmThreads) {
    const long long c = cidx[k], j = pidx[k];
    T v0 = T(0), v1 = T(0);
#pragma unroll
    for (int a = 0; a < 6; ++a) { const T xa = __ldg(x + c * 6 + a); v0 += Jc[k * 12 + a] * xa; v1 += Jc[k * 12 + 6 + a] * xa; }
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(t + j * 3 + a, Jp[k * 6 + a] * v0 + Jp[k * 6 + 3 + a] * v1);
  }
}
// y[c] += Jc_k^T (Jp_k v[j_k])      (W v, points -> cameras)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_ba_wv_kernel(const T* __restrict__ Jc, const T* __restrict__ Jp,
                                                               const int* __restrict__ cidx, const int* __restrict__ pidx,
                                                               const T* __restrict__ v, T* __restrict__ y, long long m) {
  const int lane = threadIdx.x & 31;
  for (long long k0 = (long long)blockIdx.x * kLmThreads + (threadIdx.x - lane); k0 < m; k0 += (long long)gridDim.x * kLmThreads) {
    const long long k = k0 + lane;
    const bool active = k < m;
    long long c = 0;
    T out[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    if (active) {
      c = cidx[k];
      const long long j = pidx[k];
      T u0 = T(0), u1 = T(0);
#pragma unroll
      for (int a = 0; a < 3; ++a) { const T va = __ldg(v + j * 3 + a); u0 += Jp[k * 6 + a] * va; u1 += Jp[k * 6 + 3 + a] * va; }
#pragma unroll
      for (int a = 0; a < 6; ++a) out[a] = Jc[k * 12 + a] * u0 + Jc[k * 12 + 6 + a] * u1;
    }
    seg_atomic_add<T, 6>(y + c * 6, c, out, active);
  }
}
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_ba_loss_kernel(const T* __restrict__ poses, const T* __restrict__ points,
                                                                 const T* __restrict__ pix, const int* __restrict__ cidx,
                                                                 const int* __restrict__ pidx, double* ws, int rk, T rdelta,
                                                                 long long m) {
  double acc[1] = {0.0};
  for (long long k = (long long)blockIdx.x * kLmThreads + threadIdx.x; k < m; k += (long long)gridDim.x * kLmThreads) {
    const long long c = cidx[k], j = pidx[k];
    T pr[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) pr[q] = __ldg(poses + c * 7 + q);
    const V3<T> p = mk(__ldg(points + j * 3), __ldg(points + j * 3 + 1), __ldg(points + j * 3 + 2));
    T rx, ry, rho, w;
    V3<T> y;
    reproj_residual(load_se3(pr), p, pix[k * 2], pix[k * 2 + 1], rx, ry, y);
    robust_eval(rk, rdelta, rx * rx + ry * ry, rho, w);
    acc[0] += (double)rho;
  }
  reduce_sums<1>(acc, ws);
}

// ------------------------------------------------------------------------------------------------
// Two-pose reprojection (BASELINE.json configs[4] as stated: block-sparse J^T J; SURVEY.md §8d cfg 5-full):
//   r_k = proj(T_b^-1 T_a p_k) - z_k,   proj(y) = (fx y.x / y.z + sk y.y / y.z + cx,  fy y.y / y.z + cy)
// (README.md:170-178 `project` is fx = fy = -1, sk = cx = cy = 0; function/geometry.py:60-112,171-225 point2pixel /
// reprojerr with intrinsics K is fx = K00, sk = K01, cx = K02, fy = K11, cy = K12; generalises
// examples/module/reprojpgo/reprojpgo.py:16-28).  With left perturbations and w = T_a p (world point),
// y = R_b^T (w - t_b):  d y / d xi_a = R_b^T [I, -w^],  d y / d xi_b = -R_b^T [I, -w^]  =>  J_b = -J_a =: -J.
// All residuals of one ordered pose pair (a, b) therefore add into ONE 6x6 block M = sum J^T J that enters H at
// (a,a), (b,b) and with a minus sign at (a,b), (b,a) — exactly the pose-graph edge structure, so the pairs are the
// "edges" of the block-sparse PCG (pcg.cu).  Residual rows are sorted by pair: pseg[e] .. pseg[e+1] are pair e's rows.
// ------------------------------------------------------------------------------------------------
// one warp per pose pair: M_e (21) = sum J^T J, u_e (6) = sum J^T r over the pair's rows.  sums: ws[0] = sum rho(|r|^2)
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj2_accum_kernel(const T* __restrict__ nodes, const T* __restrict__ pts,
                                                                       const T* __restrict__ pix, const int* __restrict__ pseg,
                                                                       const int* __restrict__ pa, const int* __restrict__ pb,
                                                                       Intr<T> K, T* __restrict__ M, T* __restrict__ u,
                                                                       const int* __restrict__ epos_i,
                                                                       const int* __restrict__ epos_j, T* __restrict__ Mn,
                                                                       T* __restrict__ un, double* ws, int rk, T rdelta,
                                                                       long long E) {
  const int lane = threadIdx.x & 31;
  const int wpb = kLmThreads / 32;
  double acc[1] = {0.0};
  for (long long e = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); e < E; e += (long long)gridDim.x * wpb) {
    T a7[7], b7[7];
    const long long ia = pa[e], ib = pb[e];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a7[k] = __ldg(nodes + ia * 7 + k); b7[k] = __ldg(nodes + ib * 7 + k); }
    const Elem<T> Ta = load_se3(a7), Tb = load_se3(b7);
    Acc6<T> ac;
    ac.zero();
    T loss = T(0);
    for (int k = pseg[e] + lane; k < pseg[e + 1]; k += 32) {
      const long long k0 = k;
      V3<T> w, y;
      reproj2_point(Ta, Tb, mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]), w, y);
      T rx, ry;
      reproj2_residual(K, y, pix[k0 * 2], pix[k0 * 2 + 1], rx, ry);
      T j0[6], j1[6];
      reproj2_rows(K, Tb, w, y, j0, j1);
      T rho, wt;
      robust_eval(rk, rdelta, rx * rx + ry * ry, rho, wt);
      if (rk) {
        const T sw = m_sqrt(wt);
        rx *= sw; ry *= sw;
#pragma unroll
        for (int q = 0; q < 6; ++q) { j0[q] *= sw; j1[q] *= sw; }
      }
      ac.add_row(j0, rx);
      ac.add_row(j1, ry);
      loss += rho;
    }
    Sys6<T> s = ac.finish();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        s.g[q] += __shfl_xor_sync(0xffffffffu, s.g[q], o);
#pragma unroll
        for (int bb = q; bb < 6; ++bb) s.A[q][bb] += __shfl_xor_sync(0xffffffffu, s.A[q][bb], o);
      }
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
    }
    if (lane == 0) {
      store_edge_blocks(s, e, M, u, epos_i, epos_j, Mn, un);
      acc[0] += (double)loss;
    }
  }
  reduce_sums<1>(acc, ws);
}

template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj2_loss_kernel(const T* __restrict__ nodes, const T* __restrict__ pts,
                                                                      const T* __restrict__ pix, const int* __restrict__ pseg,
                                                                      const int* __restrict__ pa, const int* __restrict__ pb,
                                                                      Intr<T> K, double* ws, int rk, T rdelta, long long E) {
  const int lane = threadIdx.x & 31;
  const int wpb = kLmThreads / 32;
  double acc[1] = {0.0};
  for (long long e = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); e < E; e += (long long)gridDim.x * wpb) {
    T a7[7], b7[7];
    const long long ia = pa[e], ib = pb[e];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a7[k] = __ldg(nodes + ia * 7 + k); b7[k] = __ldg(nodes + ib * 7 + k); }
    const Elem<T> Ta = load_se3(a7), Tb = load_se3(b7);
    T loss = T(0);
    for (int k = pseg[e] + lane; k < pseg[e + 1]; k += 32) {
      const long long k0 = k;
      V3<T> w, y;
      reproj2_point(Ta, Tb, mk(pts[k0 * 3], pts[k0 * 3 + 1], pts[k0 * 3 + 2]), w, y);
      T rx, ry, rho, wt;
      reproj2_residual(K, y, pix[k0 * 2], pix[k0 * 2 + 1], rx, ry);
      robust_eval(rk, rdelta, rx * rx + ry * ry, rho, wt);
      loss += rho;
    }
    acc[0] += (double)loss;
  }
  reduce_sums<1>(acc, ws);
}

// residual rows r (m, 2) for forward() parity; pair index per row
template <typename T>
__global__ void __launch_bounds__(kLmThreads) lm_reproj2_residual_kernel(const T* __restrict__ nodes, const T* __restrict__ pts,
                                                                          const T* __restrict__ pix, const int* __restrict__ ia,
                                                                          const int* __restrict__ ib, Intr<T> K,
                                                                          T* __restrict__ r, long long m) {
  for (long long k = (long long)blockIdx.x * kLmThreads + threadIdx.x; k < m; k += (long long)gridDim.x * kLmThreads) {
    T a7[7], b7[7];
    const long long a = ia[k], b = ib[k];
#pragma unroll
    for (int q = 0; q < 7; ++q) { a7[q] = __ldg(nodes + a * 7 + q); b7[q] = __ldg(nodes + b * 7 + q); }
    V3<T> w, y;
    reproj2_point(load_se3(a7), load_se3(b7), mk(pts[k * 3], pts[k * 3 + 1], pts[k * 3 + 2]), w, y);
    T rx, ry;
    reproj2_residual(K, y, pix[k * 2], pix[k * 2 + 1], rx, ry);
    r[k * 2] = rx;
    r[k * 2 + 1] = ry;
  }
}

}  // namespace b200pose

using namespace b200pose;

// workspace: at least b200_lm_workspace_doubles() doubles, zero-initialised ONCE by the caller (the kernels
// re-arm it themselves).  Totals are in ws[0..3] after the kernel completes.
B200_EXPORT long long b200_lm_workspace_doubles(void) { return 8 + (long long)kMaxSums * 8 * 1024; }

#define LM_ABI(SFX, CT)                                                                                               \
  B200_EXPORT int b200_lm_poseinv_loss_##SFX(const CT* P, const CT* X, double* ws, int robust, double delta,          \
                                             long long n, void* stream) {                                             \
    if (n <= 0) return 0;                                                                                             \
    lm_poseinv_loss_kernel<CT><<<lm_grid(n, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(P, X, ws, robust,     \
                                                                                                (CT)delta, n);        \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_poseinv_trial_##SFX(const CT* P, const CT* X, CT* P_trial, double* ws, double scale,        \
                                              double dmin, double dmax, int robust, double delta, long long n,        \
                                              void* stream) {                                                         \
    if (n <= 0) return 0;                                                                                             \
    lm_poseinv_trial_kernel<CT><<<lm_grid(n, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                     \
        P, X, P_trial, ws, (CT)scale, (CT)dmin, (CT)dmax, robust, (CT)delta, n);                                      \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj_accum_##SFX(const CT* poses, const CT* pts, const CT* pix, const int* seg, CT* H,    \
                                             CT* g, double* ws, int robust, double delta, long long ncam,             \
                                             void* stream) {                                                          \
    if (ncam <= 0) return 0;                                                                                          \
    lm_reproj_accum_kernel<CT><<<lm_grid(ncam, kLmThreads / 32), kLmThreads, 0, (cudaStream_t)stream>>>(              \
        poses, pts, pix, seg, H, g, ws, robust, (CT)delta, (int)ncam);                                                \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_solve6_retract_##SFX(const CT* H, const CT* g, const CT* P, CT* P_trial, CT* D, double* ws, \
                                               double scale, double dmin, double dmax, long long n, void* stream) {   \
    if (n <= 0) return 0;                                                                                             \
    lm_solve6_retract_kernel<CT><<<lm_grid(n, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                    \
        H, g, P, P_trial, D, ws, (CT)scale, (CT)dmin, (CT)dmax, n);                                                   \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj_loss_##SFX(const CT* poses, const CT* pts, const CT* pix, const int* seg,            \
                                            double* ws, int robust, double delta, long long ncam, void* stream) {     \
    if (ncam <= 0) return 0;                                                                                          \
    lm_reproj_loss_kernel<CT><<<lm_grid(ncam, kLmThreads / 32), kLmThreads, 0, (cudaStream_t)stream>>>(               \
        poses, pts, pix, seg, ws, robust, (CT)delta, (int)ncam);                                                      \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj_residual_##SFX(const CT* poses, const CT* pts, const CT* pix, const int* cidx,       \
                                                CT* r, long long m, void* stream) {                                   \
    if (m <= 0) return 0;                                                                                             \
    lm_reproj_residual_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(poses, pts, pix,   \
                                                                                                   cidx, r, m);       \
    return (int)cudaGetLastError();                                                                                   \
  }

#define PGO_ABI(SFX, CT)                                                                                              \
  B200_EXPORT int b200_lm_pgo_linearize_##SFX(const CT* nodes, const CT* Z, const int* ei, const int* ej, CT* M,      \
                                              CT* u, double* ws, int robust, double delta, long long E,               \
                                              void* stream) {                                                         \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_linearize_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                     \
        nodes, Z, ei, ej, M, u, (const int*)nullptr, (const int*)nullptr, (CT*)nullptr, (CT*)nullptr, ws, robust,     \
        (CT)delta, E);                                                                                                \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_pgo_linearize_n_##SFX(const CT* nodes, const CT* Z, const int* ei, const int* ej,           \
                                                const int* epos_i, const int* epos_j, CT* Mn, CT* un, double* ws,     \
                                                int robust, double delta, long long E, void* stream) {                \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_linearize_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                     \
        nodes, Z, ei, ej, (CT*)nullptr, (CT*)nullptr, epos_i, epos_j, Mn, un, ws, robust, (CT)delta, E);              \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_pgo_linearize_w_##SFX(const CT* nodes, const CT* Z, const int* ei, const int* ej, const CT* W, \
                                                long long w_stride, CT* M, CT* u, CT* M0, CT* u0, double* ws,          \
                                                int robust, double delta, long long E, void* stream) {                 \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_linearize_w_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                   \
        nodes, Z, ei, ej, W, w_stride, M, u, M0, u0, ws, robust, (CT)delta, E);                                       \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_pgo_scatter_##SFX(const CT* M, const CT* u, const int* ei, const int* ej, CT* Hd, CT* g,    \
                                            long long E, void* stream) {                                              \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_scatter_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(M, u, ei, ej, Hd, g,   \
                                                                                               E);                    \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_pgo_spmv_##SFX(const CT* M, const int* ei, const int* ej, const CT* x, CT* y, long long E,  \
                                         void* stream) {                                                              \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_spmv_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(M, ei, ej, x, y, E);      \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_pgo_loss_##SFX(const CT* nodes, const CT* Z, const int* ei, const int* ej, double* ws,      \
                                         int robust, double delta, long long E, void* stream) {                       \
    if (E <= 0) return 0;                                                                                             \
    lm_pgo_loss_kernel<CT><<<lm_grid(E, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(nodes, Z, ei, ej, ws,     \
                                                                                            robust, (CT)delta, E);    \
    return (int)cudaGetLastError();                                                                                   \
  }

#define BA_ABI(SFX, CT)                                                                                               \
  B200_EXPORT int b200_lm_ba_linearize_##SFX(const CT* poses, const CT* points, const CT* pix, const int* cidx,       \
                                             const int* pidx, CT* Jc, CT* Jp, CT* rs, CT* Hcc, CT* Hpp, CT* gc,       \
                                             CT* gp, double* ws, int robust, double delta, long long m,               \
                                             void* stream) {                                                          \
    if (m <= 0) return 0;                                                                                             \
    lm_ba_linearize_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                      \
        poses, points, pix, cidx, pidx, Jc, Jp, (CT*)nullptr, (const int*)nullptr, (CT*)nullptr, rs, Hcc, Hpp, gc,    \
        gp, ws, robust, (CT)delta, m);                                                                                \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_ba_wtx_##SFX(const CT* Jc, const CT* Jp, const int* cidx, const int* pidx, const CT* x,     \
                                       CT* t, long long m, void* stream) {                                            \
    if (m <= 0) return 0;                                                                                             \
    lm_ba_wtx_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(Jc, Jp, cidx, pidx, x, t,   \
                                                                                          m);                         \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_ba_wv_##SFX(const CT* Jc, const CT* Jp, const int* cidx, const int* pidx, const CT* v,      \
                                      CT* y, long long m, void* stream) {                                             \
    if (m <= 0) return 0;                                                                                             \
    lm_ba_wv_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(Jc, Jp, cidx, pidx, v, y,    \
                                                                                         m);                          \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_ba_loss_##SFX(const CT* poses, const CT* points, const CT* pix, const int* cidx,            \
                                        const int* pidx, double* ws, int robust, double delta, long long m,           \
                                        void* stream) {                                                               \
    if (m <= 0) return 0;                                                                                             \
    lm_ba_loss_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(                           \
        poses, points, pix, cidx, pidx, ws, robust, (CT)delta, m);                                                    \
    return (int)cudaGetLastError();                                                                                   \
  }

LM_ABI(f32, float)
LM_ABI(f64, double)
BA_ABI(f32, float)
BA_ABI(f64, double)
PGO_ABI(f32, float)
PGO_ABI(f64, double)

#define REPROJ2_ABI(SFX, CT)                                                                                          \
  B200_EXPORT int b200_lm_reproj2_accum_##SFX(const CT* nodes, const CT* pts, const CT* pix, const int* pseg,         \
                                              const int* pa, const int* pb, const double* intr, CT* M, CT* u,         \
                                              double* ws, int robust, double delta, long long E, void* stream) {      \
    if (E <= 0) return 0;                                                                                             \
    const Intr<CT> K = {(CT)intr[0], (CT)intr[1], (CT)intr[2], (CT)intr[3], (CT)intr[4]};                             \
    lm_reproj2_accum_kernel<CT><<<lm_grid(E, kLmThreads / 32), kLmThreads, 0, (cudaStream_t)stream>>>(                \
        nodes, pts, pix, pseg, pa, pb, K, M, u, (const int*)nullptr, (const int*)nullptr, (CT*)nullptr, (CT*)nullptr, \
        ws, robust, (CT)delta, E);                                                                                    \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj2_accum_n_##SFX(const CT* nodes, const CT* pts, const CT* pix, const int* pseg,       \
                                                const int* pa, const int* pb, const double* intr, const int* epos_i,  \
                                                const int* epos_j, CT* Mn, CT* un, double* ws, int robust,            \
                                                double delta, long long E, void* stream) {                            \
    if (E <= 0) return 0;                                                                                             \
    const Intr<CT> K = {(CT)intr[0], (CT)intr[1], (CT)intr[2], (CT)intr[3], (CT)intr[4]};                             \
    lm_reproj2_accum_kernel<CT><<<lm_grid(E, kLmThreads / 32), kLmThreads, 0, (cudaStream_t)stream>>>(                \
        nodes, pts, pix, pseg, pa, pb, K, (CT*)nullptr, (CT*)nullptr, epos_i, epos_j, Mn, un, ws, robust, (CT)delta,  \
        E);                                                                                                           \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj2_loss_##SFX(const CT* nodes, const CT* pts, const CT* pix, const int* pseg,          \
                                             const int* pa, const int* pb, const double* intr, double* ws,            \
                                             int robust, double delta, long long E, void* stream) {                   \
    if (E <= 0) return 0;                                                                                             \
    const Intr<CT> K = {(CT)intr[0], (CT)intr[1], (CT)intr[2], (CT)intr[3], (CT)intr[4]};                             \
    lm_reproj2_loss_kernel<CT><<<lm_grid(E, kLmThreads / 32), kLmThreads, 0, (cudaStream_t)stream>>>(                 \
        nodes, pts, pix, pseg, pa, pb, K, ws, robust, (CT)delta, E);                                                  \
    return (int)cudaGetLastError();                                                                                   \
  }                                                                                                                   \
  B200_EXPORT int b200_lm_reproj2_residual_##SFX(const CT* nodes, const CT* pts, const CT* pix, const int* ia,        \
                                                 const int* ib, const double* intr, CT* r, long long m,               \
                                                 void* stream) {                                                      \
    if (m <= 0) return 0;                                                                                             \
    const Intr<CT> K = {(CT)intr[0], (CT)intr[1], (CT)intr[2], (CT)intr[3], (CT)intr[4]};                             \
    lm_reproj2_residual_kernel<CT><<<lm_grid(m, kLmThreads), kLmThreads, 0, (cudaStream_t)stream>>>(nodes, pts, pix,  \
                                                                                                    ia, ib, K, r, m); \
    return (int)cudaGetLastError();                                                                                   \
  }

REPROJ2_ABI(f32, float)
REPROJ2_ABI(f64, double)
